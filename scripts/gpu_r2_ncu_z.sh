mkdir -p gpurun_out
CMD="python scripts/ztally_probe.py 100000 0"
$CMD > gpurun_out/zprobe_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'ztally_ord|zkeep_kernel|zmoments_kernel' -c 3 -o /tmp/prof_z -f $CMD > gpurun_out/ncu_z.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/zprobe_plain.log | tail -3
ncu -i /tmp/prof_z.ncu-rep --page raw --csv > gpurun_out/raw_z_r2.csv 2>/dev/null
ncu -i /tmp/prof_z.ncu-rep --page source --csv > gpurun_out/source_z_r2.csv 2>/dev/null
ls -la gpurun_out/raw_z_r2.csv gpurun_out/source_z_r2.csv; tail -5 gpurun_out/ncu_z.log
