"""Upload paths on the cfg3 shape: chunked+repack (wgs_upload_gl) vs per-slab strided DMA (wgs_upload_gl_async)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wgsassign_b200 import _lib

M, N, K = int(os.environ.get("M", 1_000_000)), 500, 10
pop_of = ((np.arange(N) * K) // N).astype(np.int32)
ctx = _lib.Context(0)
ctx.set_pops(pop_of, K)
ctx.synth(M, N, seed=1)
Lh = _lib.pinned_empty((M, 2 * N), np.float32)
for s0 in range(0, M, 100_000):
    n = min(100_000, M - s0)
    Lh[s0:s0 + n] = ctx.download(s0, n)
gb = Lh.nbytes / 1e9
for name in ("sync", "async", "sync", "async"):
    ctx.set_pops(pop_of, K)
    t0 = time.perf_counter()
    if name == "sync":
        ctx.upload_gl(Lh)
    else:
        ctx.upload_gl_async(Lh)
        t1 = time.perf_counter()
        ctx.upload_wait()
    dt = time.perf_counter() - t0
    print("%-6s %.1f ms  %.1f GB/s%s" % (name, dt * 1e3, gb / dt, "" if name == "sync" else "  (queued in %.2f ms)" % ((t1 - t0) * 1e3)))
os.environ["WGS_TRACE"] = "1"
for _ in range(2):
    ctx.set_pops(pop_of, K)
    t0 = time.perf_counter()
    ctx.upload_gl_async(Lh)
    out = ctx.ref_af_loo(200, 1e-4)
    print("fused e2e %.1f ms" % ((time.perf_counter() - t0) * 1e3))
