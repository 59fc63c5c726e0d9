mkdir -p gpurun_out
./scripts/microbench/seqsum_rate.bin | tail -4
python scripts/seqsum_probe.py | tail -2
timeout 900 python -m pytest tests/test_gpu_exact_stop.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 5 --no-cpu-baseline --no-e2e --no-extra"
WGS_DEBUG=1 WGS_RMSE_BAND_PPM=-1 $B > gpurun_out/r2m_always.json 2> gpurun_out/r2m_always.err; python scripts/bench_brief.py gpurun_out/r2m_always.json 2>/dev/null | sed -n '1,2p;8p'
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29633 tests/multi_gpu_check.py > gpurun_out/r2_multi${N}d.log 2>&1; echo "multi rc=$?"; grep "FAIL\|PASS" gpurun_out/r2_multi${N}d.log | tail -5
WGS_DEBUG=1 WGS_RMSE_BAND_PPM=-1 timeout 600 $TR --master-port 29635 bench.py --gpus $N --steps 5 --no-extra --no-e2e > gpurun_out/r2m_weak2_always.json 2> gpurun_out/r2m_weak2_always.err; python scripts/bench_brief.py gpurun_out/r2m_weak2_always.json 2>/dev/null | sed -n '1,2p;8p'
