B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
$B > gpurun_out/b0.json 2>gpurun_out/b0.err; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 1,2p; tail -1 gpurun_out/b0.err
WGS_LOO_OCC2=1 $B > gpurun_out/b1.json 2>gpurun_out/b1.err; echo occ2; python scripts/bench_brief.py gpurun_out/b1.json | sed -n 1,2p; tail -1 gpurun_out/b1.err
