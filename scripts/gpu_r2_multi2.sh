# 2-GPU check at the final commit of the round: sharded parity (tests/multi_gpu_check.py) + the weak-scaling bench line
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29633 tests/multi_gpu_check.py > gpurun_out/r2_multi${N}_final.log 2>&1; echo "multi rc=$?"; grep "FAIL\|PASS" gpurun_out/r2_multi${N}_final.log | tail -40
timeout 600 $TR --master-port 29635 bench.py --gpus $N --steps 5 --no-extra > gpurun_out/r2_weak2_final.json 2> gpurun_out/r2_weak2_final.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2_weak2_final.json 2>/dev/null | sed -n '1,3p'; python scripts/bench_brief.py gpurun_out/r2_weak2_final.json 2>/dev/null | grep "idle\|parity"
