timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r1_2gpu_v4.json 2> gpurun_out/bench_r1_2gpu_v4.err; echo "rc=$?"
python scripts/bench_brief.py gpurun_out/bench_r1_2gpu_v4.json | head -4
tail -2 gpurun_out/bench_r1_2gpu_v4.err
