"""Functional + timing probe at the shape of BASELINE.json configs[3] (2,000 individuals x 20
populations; site count scaled to one GPU's memory): --get_pop_like + --get_reference_z_score,
plus the LOO and Fisher operators.  No oracle here (parity is in tests/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wgsassign_b200 import _lib

m = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
n, k = 2000, 20
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ctx = _lib.Context(0)
pop_of = ((np.arange(n) * k) // n).astype(np.int32)
ctx.set_pops(pop_of, k)
t0 = time.perf_counter(); ctx.synth(m, n, seed=5, with_ad=True); print("synth %.2f s" % (time.perf_counter() - t0))

def timed(name, fn):
    fn()
    ctx.timing_reset(True)
    t0 = time.perf_counter(); out = fn(); dt = time.perf_counter() - t0
    line = "%-14s %8.1f ms wall" % (name, dt * 1e3)
    for fam in ("em_pop", "loo_em", "loo_like", "pop_like", "fisher", "ztally", "zkeep", "zmoments"):
        t = ctx.timing_get(fam)
        if t["launches"]:
            line += " | %s %d x %.2f ms %.0f GB/s %.2e u/s" % (fam, t["launches"], t["ms"] / t["launches"], t["bytes"] / 1e9 / (t["ms"] * 1e-3), t["units"] / (t["ms"] * 1e-3))
    ctx.timing_reset(False)
    print(line); sys.stdout.flush()
    return out

af, its = timed("ref_af", lambda: ctx.ref_af(200, 1e-4))
print("  em iters", sorted(set(int(x) for x in its)))
pl = timed("pop_like", lambda: ctx.pop_like_partial(af))
print("  self-assignment", float(np.mean(np.argmax(pl, 1) == pop_of)), " evals/s %.3e" % (m * n * k / 1.0))
fo = timed("fisher", lambda: ctx.fisher_partial(af))
ll = timed("loo", lambda: ctx.loo_partial(None, 200, 1e-4))
print("  LOO self-assignment", float(np.mean(np.argmax(ll[0], 1) == pop_of)), "iters", int(ll[2].min()), int(ll[2].max()))
z = timed("z reference", lambda: ctx.zscore(1, None, 0, False, 0, nz, 200, 1e-4))
zz = np.array([r.z for r in z]); print("  z mean %.3f sd %.3f kept %.1f%%" % (zz.mean(), zz.std(), 100 * np.mean([r.loci_kept for r in z]) / m))
z = timed("z assignment", lambda: ctx.zscore(0, af, 0, False, 0, nz, 200, 1e-4))
