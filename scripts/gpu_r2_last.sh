# last call of the round: smoke() and the driver's bench command at the final commit
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r2_last.json 2> gpurun_out/r2_last.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2_last.json 2>/dev/null | grep -v ingest
