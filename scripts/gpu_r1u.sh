timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/e2e_pipe.json 2>gpurun_out/e2e_pipe.err; python scripts/bench_brief.py gpurun_out/e2e_pipe.json | sed -n 1,4p; tail -2 gpurun_out/e2e_pipe.err
WGS_E2E_SERIAL=1 $B > gpurun_out/e2e_serial.json 2>gpurun_out/e2e_serial.err; python scripts/bench_brief.py gpurun_out/e2e_serial.json | sed -n 1,1p; tail -2 gpurun_out/e2e_serial.err
