mkdir -p gpurun_out
B="timeout 400 python bench.py --steps 5 --no-cpu-baseline --no-e2e --no-extra"
echo default; $B 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | sed -n '1p;3p'
for mb in 3 4; do
echo "minb $mb"; WGS_DEBUG=1 WGS_LOO_FIRST_BPSM=$mb WGS_B200_LIB=$PWD/build_variants/libwgs_f$mb.so $B 2>/dev/null | python scripts/bench_brief.py /dev/stdin 2>/dev/null | sed -n '1p;3p'
done
