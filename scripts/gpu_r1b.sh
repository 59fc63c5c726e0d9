set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_v3.json 2> gpurun_out/bench_r1_v3.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/bench_r1_v3.json
WGS_LOO_V2=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1_v2ab.json 2> gpurun_out/bench_r1_v2ab.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/bench_r1_v2ab.json
