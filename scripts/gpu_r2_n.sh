mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29633 tests/multi_gpu_check.py > gpurun_out/r2_multi${N}e.log 2>&1; echo "multi rc=$?"; grep "FAIL\|PASS" gpurun_out/r2_multi${N}e.log | tail -5
timeout 600 $TR --master-port 29635 bench.py --gpus $N --steps 5 --no-extra > gpurun_out/r2n_weak2.json 2> gpurun_out/r2n_weak2.err; python scripts/bench_brief.py gpurun_out/r2n_weak2.json 2>/dev/null | sed -n '1,2p;8p'
WGS_DEBUG=1 WGS_RMSE_BAND_PPM=20000 timeout 600 $TR --master-port 29636 bench.py --gpus $N --steps 5 --no-extra > gpurun_out/r2n_weak2_band2pct.json 2> gpurun_out/r2n_weak2_band2pct.err; python scripts/bench_brief.py gpurun_out/r2n_weak2_band2pct.json 2>/dev/null | sed -n '1,2p;8p'
