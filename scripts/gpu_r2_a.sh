# round 2, call A: GPU tests after the options / sequential-tally / exact-stop-rule rework + a short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/r2a_bench.json | head -12; tail -3 gpurun_out/r2a_bench.err
