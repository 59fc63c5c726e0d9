mkdir -p gpurun_out
for u in 1 2 8; do echo "unroll $u"; WGS_B200_LIB=$PWD/build_variants/libwgs_u$u.so python scripts/ztally_probe.py 1000000 1 2>&1 | tail -1; done
echo "unroll 4 (default)"; python scripts/ztally_probe.py 1000000 1 2>&1 | tail -1
