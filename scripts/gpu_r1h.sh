python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/b0.json 2>gpurun_out/b0.err; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 1,6p; tail -2 gpurun_out/b0.err
WGS_LOO_V4=1 $B --no-extra > gpurun_out/b1.json 2>gpurun_out/b1.err; echo "v4"; python scripts/bench_brief.py gpurun_out/b1.json | sed -n 1,2p
WGS_PL2_WX=3 $B > gpurun_out/b2.json 2>gpurun_out/b2.err; echo "wx 3"; python scripts/bench_brief.py gpurun_out/b2.json | sed -n 5,5p
