set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
export WGS_BENCH_ALLOW_SHORT=1
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sites 200000"
$SHORT > gpurun_out/plain_short.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
$SHORT > gpurun_out/plain_short2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:loo_em_step -s 30 -c 3 -o gpurun_out/prof_loo_em_r1 $SHORT > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/
