"""Timing of the order-exact float32 summation kernel (one block) and of the LOO resolve at cfg3 shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wgsassign_b200 import _lib
ctx = _lib.Context(0)
rng = np.random.default_rng(0)
for n in (1_000_000, 8_000_000):
    x = ((rng.standard_normal(n) * 1e-4) ** 2).astype(np.float32)
    ctx.debug_seqsum(x)
    ctx.timing_reset(True)
    for _ in range(3):
        r = ctx.debug_seqsum(x)
    t = ctx.timing_get("em_resolve")
    ref = np.cumsum(x, dtype=np.float32)[-1]
    print("n=%d: %.3f ms per sum (%.2f ns per addend), bits equal %s" % (n, t["ms"] / t["launches"], t["ms"] / t["launches"] * 1e6 / n, r.tobytes() == ref.tobytes()))
    ctx.timing_reset(False)
