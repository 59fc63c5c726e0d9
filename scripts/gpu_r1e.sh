B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
$B > gpurun_out/b0.json 2>gpurun_out/b0.err; python scripts/bench_brief.py gpurun_out/b0.json | sed -n 1,2p
WGS_LOO_PASSES=1 $B > gpurun_out/b1.json 2>gpurun_out/b1.err; echo "passes 1"; python scripts/bench_brief.py gpurun_out/b1.json | sed -n 1,2p
WGS_LOO_BLOCK=128 WGS_LOO_PASSES=2 $B > gpurun_out/b2.json 2>gpurun_out/b2.err; echo "block 128 passes 2"; python scripts/bench_brief.py gpurun_out/b2.json | sed -n 1,2p
WGS_LOO_BLOCK=128 WGS_LOO_PASSES=4 $B > gpurun_out/b3.json 2>gpurun_out/b3.err; echo "block 128 passes 4"; python scripts/bench_brief.py gpurun_out/b3.json | sed -n 1,2p
WGS_LOO_BLOCK=384 $B > gpurun_out/b4.json 2>gpurun_out/b4.err; echo "block 384"; python scripts/bench_brief.py gpurun_out/b4.json | sed -n 1,2p
export WGS_BENCH_ALLOW_SHORT=1
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra --sites 200000"
K=loo_em_step4
$SHORT > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 -o /tmp/prof_${K} -f $SHORT > gpurun_out/ncu_${K}.log 2>&1
echo "$K rc=$?"
ncu -i /tmp/prof_${K}.ncu-rep --page raw --csv > gpurun_out/raw_${K}_r1v8.csv 2>/dev/null
ncu -i /tmp/prof_${K}.ncu-rep --page source --csv > gpurun_out/source_${K}_r1v8.csv 2>/dev/null
