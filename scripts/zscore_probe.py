"""Times the z-score passes on a device-generated matrix (no oracle; parity is in tests/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wgsassign_b200 import _lib

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n, k = 500, 10
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else n
ctx = _lib.Context(0)
pop_of = ((np.arange(n) * k) // n).astype(np.int32)
ctx.set_pops(pop_of, k)
ctx.synth(m, n, seed=3, with_ad=True)
af, _ = ctx.ref_af(200, 1e-4)
for mode, name in ((0, "assignment"), (1, "reference")):
    for exact in ("0", "1"):
        os.environ["WGS_Z_EXACT_MEANS"] = exact
        ctx.zscore(mode, af if mode == 0 else None, 0, False, 0, min(nsub, 64))     # warm
        ctx.timing_reset(True)
        t0 = time.perf_counter()
        rows = ctx.zscore(mode, af if mode == 0 else None, 0, False, 0, nsub)
        dt = time.perf_counter() - t0
        z = np.array([r.z for r in rows]); kept = np.array([r.loci_kept for r in rows])
        line = "%-10s exact_means=%s  %d inds x %d sites: %.1f ms wall; z mean %.3f sd %.3f; kept %.1f%%" % (
            name, exact, nsub, m, dt * 1e3, np.nanmean(z), np.nanstd(z), 100 * kept.mean() / m)
        for fam in ("ztally", "zkeep", "zmoments", "loo_em"):
            t = ctx.timing_get(fam)
            if t["launches"]:
                line += " | %s %.2f ms" % (fam, t["ms"])
                if t["bytes"]:
                    line += " (%.0f GB/s)" % (t["bytes"] / 1e9 / (t["ms"] * 1e-3))
        print(line)
        ctx.timing_reset(False)
