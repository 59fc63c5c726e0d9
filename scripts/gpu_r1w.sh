timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/b1.json 2>gpurun_out/b1.err; python scripts/bench_brief.py gpurun_out/b1.json | sed -n 1,8p; tail -2 gpurun_out/b1.err
WGS_TRACE=1 python scripts/cfg4_probe.py 2>&1 | grep -v "loo  \|ref_af  " | tail -40
