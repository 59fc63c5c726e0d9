set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_1gpu_v2.json 2> gpurun_out/bench_r1_1gpu_v2.err; echo "bench rc=$?"
python scripts/bench_brief.py gpurun_out/bench_r1_1gpu_v2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref_v2.json 2> gpurun_out/bench_r1_ref_v2.err; echo "ref rc=$?"; cat gpurun_out/bench_r1_ref_v2.json | cut -c1-400
scripts/microbench/loo_quad_rate.bin > gpurun_out/loo_quad_rate_r1.txt; cat gpurun_out/loo_quad_rate_r1.txt
python scripts/cfg4_probe.py > gpurun_out/cfg4_probe_r1.txt 2>&1; cat gpurun_out/cfg4_probe_r1.txt
bash scripts/gpu_profile_all.sh r1f > gpurun_out/profile_all.log 2>&1; tail -3 gpurun_out/profile_all.log
