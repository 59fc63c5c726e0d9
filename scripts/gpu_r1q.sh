B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
for st in 0 3 4; do
if [ $st = 0 ]; then unset WGS_LOO_STAGES; else export WGS_LOO_STAGES=$st; fi
$B > gpurun_out/b0.json 2>gpurun_out/b0.err; echo "stages $st"; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 1,2p; tail -1 gpurun_out/b0.err
done
unset WGS_LOO_STAGES
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
