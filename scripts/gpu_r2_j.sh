# round 2, call J: full GPU suite, default bench line (with parity + ingest), ncu of the main kernels at cfg3 shape
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_pytest.log
timeout 500 python bench.py > gpurun_out/r2j_cfg3.json 2> gpurun_out/r2j_cfg3.err; echo "cfg3 rc=$?"; python scripts/bench_brief.py gpurun_out/r2j_cfg3.json; tail -3 gpurun_out/r2j_cfg3.err
export WGS_BENCH_ALLOW_SHORT=1
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra --sites 200000"
$SHORT > gpurun_out/plain_j.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'loo_em_step5|loo_like3|em_pop_multi2' -s 60 -c 12 -o /tmp/prof_j -f $SHORT > gpurun_out/ncu_j.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_j.ncu-rep --page raw --csv > gpurun_out/raw_main_r2.csv 2>/dev/null
ncu -i /tmp/prof_j.ncu-rep --page source --csv > gpurun_out/source_main_r2.csv 2>/dev/null
ls -la gpurun_out/raw_main_r2.csv gpurun_out/source_main_r2.csv | cut -c20-; tail -2 gpurun_out/ncu_j.log
