mkdir -p gpurun_out
B="python bench.py --steps 5 --no-cpu-baseline --no-e2e --no-extra"
$B > gpurun_out/r2k_base.json 2> gpurun_out/r2k_base.err; python scripts/bench_brief.py gpurun_out/r2k_base.json | head -3
WGS_DEBUG=1 WGS_LOO_VARIANT=1 $B > gpurun_out/r2k_v1.json 2> gpurun_out/r2k_v1.err; python scripts/bench_brief.py gpurun_out/r2k_v1.json | head -3
WGS_DEBUG=1 WGS_LOO_OCC3=1 $B > gpurun_out/r2k_occ3.json 2> gpurun_out/r2k_occ3.err; python scripts/bench_brief.py gpurun_out/r2k_occ3.json | head -3
WGS_DEBUG=1 WGS_TRACE=1 python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2k_trace.json 2> gpurun_out/r2k_trace.err; tail -24 gpurun_out/r2k_trace.err
WGS_DEBUG=1 WGS_RMSE_EXACT=0 $B > gpurun_out/r2k_noexact.json 2> gpurun_out/r2k_noexact.err; python scripts/bench_brief.py gpurun_out/r2k_noexact.json | head -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic or fallback or golden" 2>&1 | tail -2
