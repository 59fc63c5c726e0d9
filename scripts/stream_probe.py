import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wgsassign_b200 import _lib
n, k, m = 500, 10, 1_000_000
ctx = _lib.Context(0)
ctx.set_pops(((np.arange(n) * k) // n).astype(np.int32), k)
ctx.synth(m, n, seed=1)
for mode in (0, 1, 0, 1):
    ms, nb = ctx.debug_stream(mode)
    print("mode %d: %.3f ms  %.1f GB  -> %.0f GB/s" % (mode, ms, nb / 1e9, nb / 1e9 / (ms * 1e-3)))
