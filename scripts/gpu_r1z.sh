timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/e2e_pipe.json 2>gpurun_out/e2e_pipe.err; python scripts/bench_brief.py gpurun_out/e2e_pipe.json | sed -n 1,4p; tail -2 gpurun_out/e2e_pipe.err
M=1000000 python scripts/upload_probe.py 2>&1 | tail -9
python scripts/cfg4_probe.py 2>&1 | grep -i "^loo "
