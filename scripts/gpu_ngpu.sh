N=$1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_${N}gpu.log 2>&1; tail -2 gpurun_out/pytest_multi_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
python scripts/bench_brief.py gpurun_out/bench_${N}gpu.json 2>/dev/null | head -2
