# ncu captures of the hot kernels (one GPU).  Each ncu command is preceded by the same plain run.
# Reports are converted to CSV on the box (gpurun_out/ is capped at 64 MiB).
set -x
mkdir -p gpurun_out
export WGS_BENCH_ALLOW_SHORT=1
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sites 200000"
TAG=${1:-r1}
KERNELS=${2:-"loo_em_step5 em_pop_multi pop_like2_kernel fisher_kernel loo_like_kernel loo_prepack"}
$SHORT > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
for K in $KERNELS; do
  SKIP=0; [ $K = loo_em_step5 ] && SKIP=40; [ $K = em_pop_multi ] && SKIP=1
  $SHORT > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o /tmp/prof_${K}_${TAG} -f $SHORT > gpurun_out/ncu_${K}.log 2>&1
  echo "$K rc=$?"
  ncu -i /tmp/prof_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_${K}_${TAG}.csv 2>/dev/null
  ncu -i /tmp/prof_${K}_${TAG}.ncu-rep --page source --csv > gpurun_out/source_${K}_${TAG}.csv 2>/dev/null
done
du -sh gpurun_out
