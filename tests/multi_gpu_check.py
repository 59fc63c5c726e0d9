"""Multi-GPU parity check, launched by tests/test_gpu_multi.py (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank holds a contiguous site shard on its own GPU; rank 0 additionally runs the whole
problem on one GPU.  Sharded results must equal the single-GPU results: identical EM
iteration counts, bit-identical per-site outputs (AF, Fisher), identical z-score tallies, and
the per-individual sums equal to the rounding of their FP32 per-thread partials (the thread ->
site assignment changes with the shard size): 1e-7 relative for log-likelihoods, 1e-6 for the
Fisher per-individual means."""
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
    dist.allreduce_sum(ll); dist.allreduce_sum(llp)
    pl = ctx.pop_like_partial(af); dist.allreduce_sum(pl)
    f_obs, ne_obs, ind = ctx.fisher_partial(af); dist.allreduce_sum(ind)
    zr = ctx.zscore(1, None, 0, False, 0, 12, 200, 1e-4)
    za = ctx.zscore(0, af, 0, False, 5, 20, 200, 1e-4)
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
    dist.allreduce_sum(fll)
    faf_full = dist.gather_rows(faf)
    dist.enable(m, lo, device=torch.device("cuda", local))
    ok = True
    if rank == 0:
        dist.disable()
        os.environ["WGS_Z_EXACT_MEANS"] = "1"          # the sharded path uses the order-independent tally
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
        za1 = one.zscore(0, af1, 0, False, 5, 20, 200, 1e-4)

        one2 = _lib.Context(local)
        one2.set_pops(pop_of2, 4)
        one2.upload_gl(d2["L"])
        gaf, gits = one2.ref_af(200, 1e-4)
        gll, _, glits = one2.loo_partial(gaf.copy(), 200, 1e-4)

        def rel(a, b):
            return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
        checks = {
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
