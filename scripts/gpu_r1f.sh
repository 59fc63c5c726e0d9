python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/b0.json 2>gpurun_out/b0.err; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 1,6p
for wx in 3 5; do
WGS_PL2_WX=$wx $B > gpurun_out/b1.json 2>gpurun_out/b1.err; echo "wx $wx"; python scripts/bench_brief.py gpurun_out/b1.json | sed -n 5,5p
done
python scripts/cfg4_probe.py 2>&1 | grep -i "pop_like"
