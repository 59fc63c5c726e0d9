# Round-1 closing run on one B200: smoke, GPU tests, the default bench line, the reference arm, launch list and
# full captures of the kernels that changed since the r1h set.
TAG=${1:-r1i}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/bench_${TAG}_1gpu.json | head -12
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"
python scripts/cfg4_probe.py > gpurun_out/cfg4_probe_${TAG}.txt 2>&1; tail -12 gpurun_out/cfg4_probe_${TAG}.txt
bash scripts/gpu_profile_all.sh $TAG "loo_em_step5 loo_like2" 2>&1 | grep "rc="
