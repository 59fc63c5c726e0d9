N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_exact_stop.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 $TR --master-port 29633 tests/multi_gpu_check.py > gpurun_out/r2_multi${N}b.log 2>&1; echo "multi rc=$?"; grep "FAIL\|PASS" gpurun_out/r2_multi${N}b.log | tail -5
timeout 600 $TR --master-port 29635 bench.py --gpus $N --steps 5 --no-extra > gpurun_out/r2b_cfg3_weak_${N}gpu.json 2> gpurun_out/r2b_cfg3_weak_${N}gpu.err; echo "cfg3 weak rc=$?"; python scripts/bench_brief.py gpurun_out/r2b_cfg3_weak_${N}gpu.json 2>/dev/null | head -9
timeout 600 $TR --master-port 29634 bench.py --gpus $N --config cfg4 --sites 500000 --steps 2 --no-e2e > gpurun_out/r2b_cfg4_500k_${N}gpu.json 2> gpurun_out/r2b_cfg4_500k_${N}gpu.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2b_cfg4_500k_${N}gpu.json 2>/dev/null | sed -n 1,11p
