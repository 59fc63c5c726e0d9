mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2o_tests.log
timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/r2o_b256.json 2> gpurun_out/r2o_b256.err; python scripts/bench_brief.py gpurun_out/r2o_b256.json 2>/dev/null | grep -v ingest | head -12
WGS_DEBUG=1 WGS_LOO_BLOCK=320 timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/r2o_b320.json 2> gpurun_out/r2o_b320.err; python scripts/bench_brief.py gpurun_out/r2o_b320.json 2>/dev/null | grep -v ingest | head -12
timeout 400 python bench.py --config cfg5 --sites 400000 --steps 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2o_c5_256.json 2> gpurun_out/r2o_c5_256.err; python scripts/bench_brief.py gpurun_out/r2o_c5_256.json 2>/dev/null | grep -v ingest | head -12
WGS_DEBUG=1 WGS_LOO_BLOCK=320 timeout 400 python bench.py --config cfg5 --sites 400000 --steps 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2o_c5_320.json 2> gpurun_out/r2o_c5_320.err; python scripts/bench_brief.py gpurun_out/r2o_c5_320.json 2>/dev/null | grep -v ingest | head -12
