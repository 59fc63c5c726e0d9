"""Multi-GPU parity check, launched by tests/test_gpu_multi.py (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank holds a contiguous site shard on its own GPU; rank 0 additionally runs the whole
problem on one GPU AND through the CPU oracle.  Sharded results must equal the single-GPU
results: identical EM iteration counts, bit-identical per-site outputs (AF, Fisher), and the
per-individual sums equal to the rounding of their FP32 per-thread partials (the thread ->
site assignment changes with the shard size): 1e-7 relative for log-likelihoods, 1e-6 for the
Fisher per-individual means.  The sharded z-scores (both modes) are compared with the ORACLE:
loci kept, class tallies (AD_array) and EM iterations bit-exact, W_obs / z_mu / z_var within
1e-6 relative - the class means are the reference's sequential float32 sums, handed from rank
to rank in site order."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

from wgsassign_b200 import _lib, dist, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    m, n, k = 40000, 60, 4
    d = synth.synth(m, n, k, seed=77, interleave=True)
    L, AD, IDs = d["L"], d["AD"], d["IDs"]
    pops, pop_of = np.unique(IDs[:, 1], return_inverse=True)
    pop_of = pop_of.astype(np.int32)
    lo, hi = dist.shard_range(m, rank, world)
    dist.enable(m, lo, device=torch.device("cuda", local))
    ctx = _lib.Context(local)
    ctx.set_pops(pop_of, k)
    ctx.upload_gl(np.ascontiguousarray(L[lo:hi]))
    ctx.upload_ad(np.ascontiguousarray(AD[lo:hi]))
    dist.attach(ctx)
    af, its = ctx.ref_af(200, 1e-4)
    af_in = af.copy()
    ll, llp, lits = ctx.loo_partial(af_in, 200, 1e-4, parts=2)
    dist.combine(ctx, ll); dist.combine(ctx, llp)
    pl = ctx.pop_like_partial(af); dist.combine(ctx, pl)
    f_obs, ne_obs, ind = ctx.fisher_partial(af); dist.combine(ctx, ind)
    zr = ctx.zscore(1, None, 0, False, 0, 12, 200, 1e-4)
    zr_cls = [ctx.zscore_classes(i) for i in range(0, 12)]
    za = ctx.zscore(0, af, 0, False, 5, 20, 200, 1e-4)
    za_cls = [ctx.zscore_classes(i) for i in range(5, 20)]
    combined = ctx.partials_combined()
    # every stop check resolved with the chained sequential float32 sum (exercises the rank chain and, under NCCL, the
    # restart when a check needs it before the chain is queued): the same decisions and states as the banded default
    ctx.set_option("rmse_band_ppm", -1)
    afx, itsx = ctx.ref_af(200, 1e-4)
    afx_in = afx.copy()
    llx, _, litsx = ctx.loo_partial(afx_in, 200, 1e-4)
    ctx.set_option("rmse_band_ppm", 0)
    chain_same = list(itsx) == list(its) and list(litsx) == list(lits) and np.array_equal(afx, af) and np.array_equal(afx_in, af_in)
    # the same shard without a communicator: stop rule, table hand-over and sums through the host callback
    ctxh = _lib.Context(local)
    ctxh.set_pops(pop_of, k)
    ctxh.upload_gl(np.ascontiguousarray(L[lo:hi]))
    ctxh.upload_ad(np.ascontiguousarray(AD[lo:hi]))
    dist.attach(ctxh, nccl=False)
    zrh = ctxh.zscore(1, None, 0, False, 0, 6, 200, 1e-4)
    llh, _, litsh = ctxh.loo_partial(af.copy(), 200, 1e-4)
    dist.combine(ctxh, llh)
    host_same = (not ctxh.partials_combined()
                 and all((a.loci_kept, a.em_iters, a.z, a.w_obs, a.z_mu, a.z_var) == (b.loci_kept, b.em_iters, b.z, b.w_obs, b.z_mu, b.z_var)
                         for a, b in zip(zrh, zr[:6]))
                 and list(litsh) == list(lits) and np.array_equal(llh, ll))
    ctxh.close()
    af_full, af_loo_full, f_full = dist.gather_rows(af), dist.gather_rows(af_in), dist.gather_rows(f_obs)
    # the fused, pipelined call (asynchronous slab upload + wgs_ref_af_loo) on a second, population-contiguous data set
    m2 = 30000
    d2 = synth.synth(m2, 44, 4, seed=78, interleave=False, with_ad=False)
    pops2, pop_of2 = np.unique(d2["IDs"][:, 1], return_inverse=True)
    pop_of2 = pop_of2.astype(np.int32)
    lo2, hi2 = dist.shard_range(m2, rank, world)
    dist.enable(m2, lo2, device=torch.device("cuda", local))
    L2 = _lib.pinned_empty((hi2 - lo2, d2["L"].shape[1]), np.float32)
    L2[...] = d2["L"][lo2:hi2]
    ctx2 = _lib.Context(local)
    ctx2.set_pops(pop_of2, 4)
    dist.attach(ctx2)
    ctx2.upload_gl_async(L2)
    faf, fits, fll, _, flits, _ = ctx2.ref_af_loo(200, 1e-4)
    dist.combine(ctx2, fll)
    faf_full = dist.gather_rows(faf)
    dist.enable(m, lo, device=torch.device("cuda", local))
    ok = True
    if rank == 0:
        dist.disable()
        one = _lib.Context(local)
        one.set_pops(pop_of, k)
        one.upload_gl(L)
        one.upload_ad(AD)
        af1, its1 = one.ref_af(200, 1e-4)
        a1 = af1.copy()
        ll1, llp1, lits1 = one.loo_partial(a1, 200, 1e-4, parts=2)
        pl1 = one.pop_like_partial(af1)
        f1, ne1, ind1 = one.fisher_partial(af1)
        zr1 = one.zscore(1, None, 0, False, 0, 12, 200, 1e-4)
        zr1_cls = [one.zscore_classes(i) for i in range(0, 12)]
        za1 = one.zscore(0, af1, 0, False, 5, 20, 200, 1e-4)

        one2 = _lib.Context(local)
        one2.set_pops(pop_of2, 4)
        one2.upload_gl(d2["L"])
        gaf, gits = one2.ref_af(200, 1e-4)
        gll, _, glits = one2.loo_partial(gaf.copy(), 200, 1e-4)

        def rel(a, b):
            return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

        # the oracle on the whole matrix (test infrastructure): what the reference computes for these individuals
        from oracle import oracle

        def sort_rows(a):
            a = np.asarray(a).reshape(-1, 4)
            return a[np.lexsort((a[:, 1], a[:, 2]))]

        def z_vs_oracle(rows, classes, ref):
            good = True
            for r, cl, o in zip(rows, classes, ref):
                good = good and r.loci_kept == o["loci_kept"] and np.array_equal(sort_rows(cl), sort_rows(o["AD_array"]))
                for a, b in ((r.w_obs, o["w_obs"]), (r.z_mu, o["z_mu"]), (r.z_var, o["z_var"])):
                    good = good and abs(float(a) - float(b)) <= 1e-6 * abs(float(b))
                tol = 2e-6 * (abs(float(o["w_obs"])) + abs(float(o["z_mu"]))) / np.sqrt(float(o["z_var"])) + 1e-5 * abs(float(o["z"]))
                good = good and abs(float(r.z) - float(o["z"])) <= tol
            return bool(good)
        threads = os.cpu_count() or 1
        zr_o = oracle.zscore_reference(L, AD, IDs, 200, 1e-4, ind_start=0, ind_end=12, t=threads)
        za_o = oracle.zscore_assignment(L, AD, af1, IDs, pops, ind_start=5, ind_end=20, t=threads)
        checks = {
            "nccl sums": combined,
            "all checks sequential == banded": chain_same,
            "host-callback path == nccl path": host_same,
            "z ref == oracle": z_vs_oracle(zr, zr_cls, zr_o) and [r.em_iters for r in zr] == [o["em_iter"] for o in zr_o],
            "z asg == oracle": z_vs_oracle(za, za_cls, za_o),
            "z 1gpu == oracle": z_vs_oracle(zr1, zr1_cls, zr_o),
            "fused iters": list(fits) == list(gits) and list(flits) == list(glits),
            "fused af bitwise": np.array_equal(faf_full, gaf),
            "fused ll": rel(fll, gll) < 1e-7,
            "ref_af iters": list(its) == list(its1),
            "ref_af bitwise": np.array_equal(af_full, af1),
            "loo iters": list(lits) == list(lits1),
            "loo af bitwise": np.array_equal(af_loo_full, a1),
            "loo ll": rel(ll, ll1) < 1e-7,
            "loo parts": rel(llp, llp1) < 1e-7,
            "pop_like": rel(pl, pl1) < 1e-7,
            "fisher bitwise": np.array_equal(f_full, f1),
            "ne_ind": rel(ind, ind1) < 1e-6,
            "z ref kept": [r.loci_kept for r in zr] == [r.loci_kept for r in zr1],
            "z ref iters": [r.em_iters for r in zr] == [r.em_iters for r in zr1],
            "z ref comps": all(abs(a.z_mu - b.z_mu) <= 1e-6 * abs(b.z_mu) and abs(a.w_obs - b.w_obs) <= 1e-6 * abs(b.w_obs)
                               and abs(a.z_var - b.z_var) <= 1e-6 * abs(b.z_var) for a, b in zip(zr, zr1)),
            "z asg kept": [r.loci_kept for r in za] == [r.loci_kept for r in za1],
            "z asg comps": all(abs(a.z_mu - b.z_mu) <= 1e-6 * abs(b.z_mu) for a, b in zip(za, za1)),
        }
        print("rel: loo %.2e parts %.2e pop_like %.2e ne_ind %.2e" % (rel(ll, ll1), rel(llp, llp1), rel(pl, pl1), rel(ind, ind1)))
        for name, good in checks.items():
            print("%-16s %s" % (name, "ok" if good else "FAIL"))
            ok = ok and good
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", "world", world)
    td.barrier()
    td.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
