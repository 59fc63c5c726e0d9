mkdir -p gpurun_out
B="timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-e2e --no-extra"
for p in 1 2; do
$B > gpurun_out/r2p_h$p.json 2> gpurun_out/r2p_h$p.err; python scripts/bench_brief.py gpurun_out/r2p_h$p.json 2>/dev/null | grep "value\|idle"
done
timeout 900 python bench.py > gpurun_out/r2p_full.json 2> gpurun_out/r2p_full.err; echo "full rc=$?"; python scripts/bench_brief.py gpurun_out/r2p_full.json 2>/dev/null
