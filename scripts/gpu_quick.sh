timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
$B > gpurun_out/b0.json 2>gpurun_out/b0.err; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 1,5p; tail -1 gpurun_out/b0.err
python scripts/cfg4_probe.py 2>&1 | grep -i "^loo \|^z ref"
