"""Drop-in for the reference's ``WGSassign/glassy.py`` (assignLL :18-44, loo :47-112)."""
import numpy as np

from . import dist, session


def assignLL(L, af, t):
    """Log-likelihood of every individual under every population: float32 [N, K]
    (glassy.py:18-44 over glassy_cy.pyx:12-21).  Sums are float64, rounded once."""
    n, k = L.shape[1] // 2, af.shape[1]
    print(str(n) + " individuals to assign to " + str(k) + " populations")
    ctx = session.context(L)
    part = ctx.pop_like_partial(np.ascontiguousarray(af, dtype=np.float32))
    dist.combine(ctx, part)
    return part.astype(np.float32)


def loo(L, af, IDs, t, maf_iter, maf_tole, downsampled_L=None, num_partitions=1):
    """Leave-one-out assignment (glassy.py:47-112).  Returns (logl_mat [N,K] float32,
    logl_parts_mat [N*parts,K] float32) and leaves ``af`` as the reference leaves it
    (each column overwritten by the LOO estimate of that population's last individual)."""
    n, k = L.shape[1] // 2, af.shape[1]
    print(str(n) + " individuals to assign to " + str(k) + " populations")
    if downsampled_L is not None:
        print("Using downsampled GLs for likelihood evaluation in LOO assignment.")
    pop_of_ind, pops = session.pops_from_ids(IDs)
    ctx = session.context(L, pop_of_ind, len(pops))
    if downsampled_L is not None:
        session.with_downsampled(ctx, np.ascontiguousarray(downsampled_L, dtype=np.float32))
    af_work = np.ascontiguousarray(af, dtype=np.float32)
    ll, llp, its = ctx.loo_partial(af_work, maf_iter, maf_tole, use_ds=downsampled_L is not None,
                                   parts=num_partitions)
    for it in its:
        if it > 0:
            print("EM (MAF) converged at iteration: " + str(int(it)))
    dist.combine(ctx, ll)
    dist.combine(ctx, llp)
    af[...] = af_work                                   # in-place side effect (glassy.py:89)
    return ll.astype(np.float32), llp.astype(np.float32)


def loo_fused(ctx, IDs, maf_iter, maf_tole, downsampled=False, num_partitions=1):
    """`--get_reference_af --loo` as ONE device call (`Context.ref_af_loo`): the per-population EM of
    WGSassign.py:225-242 and glassy.loo (glassy.py:47-112) on a context whose matrix may still be
    uploading.  Returns (af float32 [M,K] as saved to .pop_af.npy, em iterations [K],
    logl_mat, logl_parts_mat float32, loo iterations [N]); the caller prints the messages in the
    reference's order."""
    af, its, ll, llp, lits, _ = ctx.ref_af_loo(maf_iter, maf_tole, use_ds=downsampled, parts=num_partitions)
    dist.combine(ctx, ll)
    dist.combine(ctx, llp)
    return af, its, ll.astype(np.float32), llp.astype(np.float32), lits
