# round 2, final call on one GPU: full GPU suite, launch list and `ncu --set full` captures of the step's kernels at
# 200k sites x 500 x 10 (each after the same plain run), the z-score kernels at 200k x 2,000 x 20
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_pytest.log
export WGS_BENCH_ALLOW_SHORT=1
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra --sites 200000"
TAG=r2
$SHORT > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}_final_200k.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
for K in loo_em_step5 loo_like3 loo_prepack2 loo_first_kernel em_pop_multi em_resolve; do
  SKIP=0; [ $K = loo_em_step5 ] && SKIP=40; [ $K = em_pop_multi ] && SKIP=1; [ $K = em_resolve ] && SKIP=5
  $SHORT > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o /tmp/prof_${K}_${TAG} -f $SHORT > gpurun_out/ncu_${K}.log 2>&1
  echo "$K rc=$?"
  ncu -i /tmp/prof_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_${K}_${TAG}.csv 2>/dev/null
done
ncu -i /tmp/prof_loo_em_step5_${TAG}.ncu-rep --page source --csv > gpurun_out/source_loo_em_step5_${TAG}.csv 2>/dev/null
CMD="python scripts/ztally_probe.py 200000 1"
$CMD > gpurun_out/zprobe_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'ztally_ord|zkeep_kernel|zmoments_kernel' -c 3 -o /tmp/prof_z -f $CMD > gpurun_out/ncu_z.log 2>&1
echo "z ncu rc=$?"; tail -1 gpurun_out/zprobe_plain.log
ncu -i /tmp/prof_z.ncu-rep --page raw --csv > gpurun_out/raw_z_${TAG}_final.csv 2>/dev/null
du -sh gpurun_out
