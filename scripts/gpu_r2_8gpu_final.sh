# 8-GPU lines at the final commit of the round (resident step only: no CPU baseline, no end-to-end sample, no extras)
N=8; TAG=r2final
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29634 bench.py --gpus $N --config cfg4 --steps 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/${TAG}_cfg4_${N}gpu.json 2> gpurun_out/${TAG}_cfg4_${N}gpu.err; echo "cfg4 rc=$?"; python scripts/bench_brief.py gpurun_out/${TAG}_cfg4_${N}gpu.json 2>/dev/null | head -14; tail -2 gpurun_out/${TAG}_cfg4_${N}gpu.err
timeout 90 $TR --master-port 29635 bench.py --gpus $N --steps 5 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_cfg3_weak_${N}gpu.json 2> gpurun_out/${TAG}_cfg3_weak_${N}gpu.err; echo "cfg3 weak rc=$?"; python scripts/bench_brief.py gpurun_out/${TAG}_cfg3_weak_${N}gpu.json 2>/dev/null | sed -n '1,2p'
timeout 90 $TR --master-port 29636 bench.py --gpus $N --steps 5 --scaling strong --no-extra --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_cfg3_strong_${N}gpu.json 2> gpurun_out/${TAG}_cfg3_strong_${N}gpu.err; echo "cfg3 strong rc=$?"; python scripts/bench_brief.py gpurun_out/${TAG}_cfg3_strong_${N}gpu.json 2>/dev/null | sed -n '1,2p'
