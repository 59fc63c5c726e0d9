# round 2, call C: tests + the three bench configurations on one GPU (cfg4 / cfg5 at a reduced site count)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
timeout 400 python bench.py > gpurun_out/r2c_cfg3.json 2> gpurun_out/r2c_cfg3.err; echo "cfg3 rc=$?"; python scripts/bench_brief.py gpurun_out/r2c_cfg3.json; tail -3 gpurun_out/r2c_cfg3.err
timeout 600 python bench.py --config cfg4 --sites 500000 --steps 2 > gpurun_out/r2c_cfg4_500k.json 2> gpurun_out/r2c_cfg4_500k.err; echo "cfg4 rc=$?"; python scripts/bench_brief.py gpurun_out/r2c_cfg4_500k.json; tail -3 gpurun_out/r2c_cfg4_500k.err
timeout 600 python bench.py --config cfg5 --sites 1000000 --steps 2 > gpurun_out/r2c_cfg5_1m.json 2> gpurun_out/r2c_cfg5_1m.err; echo "cfg5 rc=$?"; python scripts/bench_brief.py gpurun_out/r2c_cfg5_1m.json; tail -3 gpurun_out/r2c_cfg5_1m.err
