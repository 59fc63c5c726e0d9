B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
for d in 0 1 2; do
WGS_LOO_DBG=$d $B > gpurun_out/b0.json 2>gpurun_out/b0.err; echo "dbg $d"; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 2,2p; tail -1 gpurun_out/b0.err
done
