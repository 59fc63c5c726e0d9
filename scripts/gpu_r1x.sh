timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/b2.json 2>gpurun_out/b2.err; python scripts/bench_brief.py gpurun_out/b2.json | sed -n 5,6p; tail -2 gpurun_out/b2.err
python scripts/cfg4_probe.py 2>&1 | grep -i "^fisher\|^z "
