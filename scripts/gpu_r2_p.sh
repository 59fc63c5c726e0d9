mkdir -p gpurun_out
B="timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-e2e --no-extra"
C4="timeout 400 python bench.py --config cfg4 --sites 200000 --steps 2 --no-cpu-baseline --no-e2e --no-extra"
for rep in 1 2; do
echo "prev-slot refill"; $B 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | sed -n '1,2p'
echo "current-slot refill"; WGS_B200_LIB=$PWD/build_variants/libwgs_cur.so $B 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | sed -n '1,2p'
done
echo "prev-slot refill"; $C4 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | grep "loo_em "
echo "current-slot refill"; WGS_B200_LIB=$PWD/build_variants/libwgs_cur.so $C4 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | grep "loo_em "
