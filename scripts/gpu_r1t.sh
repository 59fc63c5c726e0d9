export WGS_BENCH_ALLOW_SHORT=1
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra --sites 200000"
K=loo_like2
$SHORT > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -o /tmp/prof_${K} -f $SHORT > gpurun_out/ncu_${K}.log 2>&1
echo "$K rc=$?"
ncu -i /tmp/prof_${K}.ncu-rep --page raw --csv > gpurun_out/raw_${K}_r1x.csv 2>/dev/null
ncu -i /tmp/prof_${K}.ncu-rep --page source --csv > gpurun_out/source_${K}_r1x.csv 2>/dev/null
