mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
B="timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-e2e --no-extra"
$B 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | sed -n '1,2p;7p'
WGS_DEBUG=1 WGS_LOO_BLOCK=256 $B 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | sed -n '1,2p'
