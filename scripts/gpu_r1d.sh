set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1_v4.json 2> gpurun_out/bench_r1_v4.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/bench_r1_v4.json
tail -3 gpurun_out/bench_r1_v4.err
