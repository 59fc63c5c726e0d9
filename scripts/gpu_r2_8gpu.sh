# 8-GPU call: sharded parity check, the cfg4 target run, cfg3 weak + strong scaling lines
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29633 tests/multi_gpu_check.py > gpurun_out/r2_multi${N}.log 2>&1; echo "multi rc=$?"; grep "ok$\|FAIL\|PASS\|rel:" gpurun_out/r2_multi${N}.log | tail -40
timeout 900 $TR --master-port 29634 bench.py --gpus $N --config cfg4 --steps 3 > gpurun_out/r2_cfg4_${N}gpu.json 2> gpurun_out/r2_cfg4_${N}gpu.err; echo "cfg4 rc=$?"; python scripts/bench_brief.py gpurun_out/r2_cfg4_${N}gpu.json 2>/dev/null | head -24; tail -3 gpurun_out/r2_cfg4_${N}gpu.err
timeout 600 $TR --master-port 29635 bench.py --gpus $N --steps 5 --no-extra > gpurun_out/r2_cfg3_weak_${N}gpu.json 2> gpurun_out/r2_cfg3_weak_${N}gpu.err; echo "cfg3 weak rc=$?"; python scripts/bench_brief.py gpurun_out/r2_cfg3_weak_${N}gpu.json 2>/dev/null | head -3
timeout 600 $TR --master-port 29636 bench.py --gpus $N --steps 5 --scaling strong --no-extra > gpurun_out/r2_cfg3_strong_${N}gpu.json 2> gpurun_out/r2_cfg3_strong_${N}gpu.err; echo "cfg3 strong rc=$?"; python scripts/bench_brief.py gpurun_out/r2_cfg3_strong_${N}gpu.json 2>/dev/null | head -3
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -3
