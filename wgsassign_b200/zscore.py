"""Drop-in for the z-score part of the reference (zscore.py:11-120 + the per-individual
loops of WGSassign.py:346-381 and :425-443).

The reference runs, per individual, two pure-Python O(M) passes (`AD_summary`,
`get_L_keep`) and two Cython kernels.  Here all individuals are processed at once by three
streaming GPU passes (`ztally`, `zkeep`, `zmoments`, see csrc/wgs_zscore.cuh); reference
mode additionally runs the batched leave-one-out EM restricted to each individual's kept
sites.  `zscore_all` is what the CLI calls.
"""
import numpy as np

from . import _lib, session

MODE_ASSIGNMENT, MODE_REFERENCE = 0, 1


def _raise_like_reference(err):
    msg = str(err)
    if "loci were kept" in msg:
        raise AssertionError(msg)                     # zscore.py:34-35 are asserts
    raise err


def zscore_all(L, AD, IDs, mode, A=None, pops=None, n_threshold=0, single_read=False,
               ind_start=0, ind_end=None, maf_iter=200, maf_tole=1e-4):
    """z-scores of individuals ``ind_start <= i < ind_end``.

    mode 0 (``--get_assignment_z_score``): ``A`` [M,K] float32 and ``pops`` (the
    ``--pop_names`` order of A's columns); individual i is scored against the column of its
    ID-file population (WGSassign.py:426-431).
    mode 1 (``--get_reference_z_score``): leave-one-out EM on the kept sites of each
    individual (WGSassign.py:352-365).

    Returns a list of dicts with z, w_obs, z_mu, z_var (float32), loci_kept, n_classes,
    em_iters - the quantities the reference prints per individual (WGSassign.py:372-380).
    """
    n = L.shape[1] // 2
    ind_end = n if ind_end is None else ind_end
    if mode == MODE_ASSIGNMENT:
        if A is None or pops is None:
            raise ValueError("assignment mode needs the AF matrix and the population names")
        pops = np.atleast_1d(np.asarray(pops))
        pop_of = np.empty(n, np.int32)
        for i in range(n):
            hit = np.argwhere(pops == IDs[i, 1])
            if hit.shape[0] == 0:
                if ind_start <= i < ind_end:
                    raise IndexError("population %r of individual %d is not in the population names" % (IDs[i, 1], i))
                pop_of[i] = 0
            else:
                pop_of[i] = hit[0][0]
        K = len(pops)
        A = np.ascontiguousarray(A, dtype=np.float32)
        if A.shape[1] != K:
            raise ValueError("AF matrix has %d columns, population names %d" % (A.shape[1], K))
    else:
        pop_of, upops = session.pops_from_ids(IDs)
        K = len(upops)
        A = None
    ctx = session.context(L, pop_of, K)
    session.with_ad(ctx, AD)
    try:
        rows = ctx.zscore(mode, A, n_threshold or 0, single_read, ind_start, ind_end, maf_iter, maf_tole)
    except _lib.WgsError as err:
        _raise_like_reference(err)
    out = []
    for j, r in enumerate(rows):
        out.append(dict(z=np.float32(r.z), w_obs=np.float32(r.w_obs), z_mu=np.float32(r.z_mu), z_var=np.float32(r.z_var),
                        loci_kept=int(r.loci_kept), n_classes=int(r.n_classes), em_iters=int(r.em_iters),
                        AD_array=ctx.zscore_classes(ind_start + j)))
    return out
