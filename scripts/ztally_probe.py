"""Probe of the z-score preparation passes (class tally, keep mask) at the cfg4 shape: used under ncu."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wgsassign_b200 import _lib
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
n, k = 2000, 20
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = _lib.Context(0)
ctx.set_pops(((np.arange(n) * k) // n).astype(np.int32), k)
ctx.synth(m, n, seed=5, with_ad=True)
af = None
if mode == 0:
    af, _ = ctx.ref_af(200, 1e-4)
for rep in range(2):
    ctx.timing_reset(True)
    t0 = time.perf_counter()
    z = ctx.zscore(mode, af, 0, False, 0, n, 200, 1e-4)
    dt = time.perf_counter() - t0
    print("rep %d: %.1f ms wall;" % (rep, dt * 1e3), {f: round(ctx.timing_get(f)["ms"], 3) for f in ("ztally", "zkeep", "zmoments")}, "kept %.3f" % (np.mean([r.loci_kept for r in z]) / m))
