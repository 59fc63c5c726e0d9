timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/b0.json 2>gpurun_out/b0.err; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 1,6p; tail -2 gpurun_out/b0.err
