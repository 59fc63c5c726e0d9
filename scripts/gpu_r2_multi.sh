# multi-GPU parity check + sharded cfg4 probe; N = number of GPUs of this call
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zscore.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 tests/multi_gpu_check.py > gpurun_out/r2_multi${N}.log 2>&1; echo "multi rc=$?"; grep -v "^\[W\|^W1\|warn" gpurun_out/r2_multi${N}.log | tail -45
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29634 bench.py --gpus $N --config cfg4 --sites 500000 --steps 2 --no-e2e > gpurun_out/r2_cfg4_500k_${N}gpu.json 2> gpurun_out/r2_cfg4_500k_${N}gpu.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2_cfg4_500k_${N}gpu.json 2>/dev/null | sed -n 1,10p; tail -3 gpurun_out/r2_cfg4_500k_${N}gpu.err
