TAG=${1:-r1j}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/bench_${TAG}_1gpu.json | head -9
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"
bash scripts/gpu_profile_all.sh $TAG "em_pop_multi2" 2>&1 | grep "rc="
