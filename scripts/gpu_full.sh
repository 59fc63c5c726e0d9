# Round-1 final measurement set on one B200: GPU tests, the default bench line, the reference arm, the ncu launch
# list + one full capture per hot kernel, the cfg4-shape probe.  Everything lands in gpurun_out/.
TAG=${1:-r1h}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/bench_${TAG}_1gpu.json | head -12
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_${TAG}_ref.json | cut -c1-600
python scripts/cfg4_probe.py > gpurun_out/cfg4_probe_${TAG}.txt 2>&1; cat gpurun_out/cfg4_probe_${TAG}.txt
bash scripts/gpu_profile_all.sh $TAG "loo_em_step5 loo_first em_pop_multi2 pop_like2_kernel fisher2 loo_like2 loo_prepack2" 2>&1 | grep "rc="
