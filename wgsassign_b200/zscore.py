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


# ---------------------------------------------------------------------------------------
# The reference's per-individual functions (zscore.py:11-120), same names, argument order and
# return types, so WGSassign.py:346-381 / :425-443 run unchanged on top of them.  Every O(M)
# pass (class tally, keep test, the two moment kernels) runs on the GPU over the resident
# matrices; only the class bookkeeping over a few dozen rows stays in NumPy, as it is in the
# reference.  `zscore_all` above does the same work for all individuals in three launches and
# is what the CLI uses.
# ---------------------------------------------------------------------------------------
def _flat_ctx(L, AD):
    ctx = session.context(L, None, 0)
    session.with_ad(ctx, AD)
    return ctx


def AD_summary(L, AD, i, n_threshold, single_read_threshold):
    """zscore.AD_summary (zscore.py:11-41): ``(AD_summary_dict, AD_array)``.

    ``AD_summary_dict[(ref, alt)] = [n_loci, float32 mean GL triple]`` with keys in
    first-occurrence order (the order of the reference's dict), ``AD_array`` int32 ``[C,4]``
    = ``[a1, a2, a1+a2, n_loci]`` after the same filters.  Tallies and means come from the
    ``ztally`` kernel; the first-occurrence order is read off the caller's AD column."""
    ctx = _flat_ctx(L, AD)
    try:
        ctx.zscore(2, None, n_threshold or 0, single_read_threshold, i, i + 1)
    except _lib.WgsError as err:
        _raise_like_reference(err)
    tab = ctx.zscore_table(i)
    ref_col = np.ascontiguousarray(AD[:, 2 * i]).astype(np.int64)
    alt_col = np.ascontiguousarray(AD[:, 2 * i + 1]).astype(np.int64)
    base = int(alt_col.max()) + 1 if alt_col.size else 1
    keys, first = np.unique(ref_col * base + alt_col, return_index=True)
    order = {(int(k // base), int(k % base)): int(f) for k, f in zip(keys, first)}
    rows = sorted(tab, key=lambda r: order[(int(r[0]), int(r[1]))])
    AD_summary_dict = {}
    for r in rows:
        AD_summary_dict[(np.int32(r[0]), np.int32(r[1]))] = [int(r[2]), np.asarray(r[3:6], dtype=np.float32)]
    arr = np.array([[int(r[0]), int(r[1]), int(r[0]) + int(r[1]), int(r[2])] for r in rows], dtype=np.int32).reshape(-1, 4)
    if single_read_threshold:
        AD_filtered = arr[arr[:, 2] == 1]
    else:
        AD_filtered = arr[(arr[:, 3] > n_threshold) & (arr[:, 2] != 0)]
    assert (AD_filtered.shape[0] != 0), "No loci were kept! Too stringent filtering?"
    assert (AD_filtered.shape[0] != 1), "Not enough loci were kept! Too stringent filtering?"
    dl, dl_counts = np.unique(AD_filtered[:, 0] + AD_filtered[:, 1], return_counts=True)
    dl_keep = dl[dl < dl_counts]
    AD_array = AD_filtered[np.isin(AD_filtered[:, 2], dl_keep)]
    return AD_summary_dict, AD_array


def get_L_keep(L, AD, AD_summary_dict, AD_array, i):
    """zscore.get_L_keep (zscore.py:43-61): ``(L_keep int32 [M_keep], loci_kept)`` - the sites
    whose (ref, alt) class is in AD_array and whose GL at the class's arg-max genotype is
    within 0.01 of the class mean (``zkeep`` kernel)."""
    ctx = _flat_ctx(L, AD)
    means = np.zeros((AD_array.shape[0], 3), np.float32)
    for c in range(AD_array.shape[0]):
        means[c] = AD_summary_dict[(AD_array[c, 0], AD_array[c, 1])][1]
    keep = ctx.zkeep_one(i, AD_array, means)
    return keep, keep.shape[0]


def get_factorials(AD_array, AD_summary_dict, e):
    """zscore.get_factorials (zscore.py:63-79): multinomial read-count probabilities, class
    mean GLs and the (ref, alt) -> class-row table.  A few dozen rows of host arithmetic in
    float64 stored as float32, exactly as the reference (``math.factorial`` replaces the
    ``np.math`` alias numpy removed)."""
    import math
    C = AD_array.shape[0]
    AD_factorial = np.zeros((C, 3), dtype=np.float32)
    AD_like = np.zeros((C, 3), dtype=np.float32)
    AD_index = np.zeros((np.max(AD_array[:, 0]) + 1, np.max(AD_array[:, 1]) + 1), dtype=np.int32)
    for c in range(C):
        Ar, Aa = int(AD_array[c, 0]), int(AD_array[c, 1])
        AD_index[Ar, Aa] = np.argwhere((AD_array[:, 0] == Ar) & (AD_array[:, 1] == Aa))[0][0]
        Dl = Aa + Ar
        coef = math.factorial(Dl) / (math.factorial(Aa) * math.factorial(Ar))
        AD_factorial[c, :] = [coef * ((1.0 - e) ** Ar) * (e ** Aa), coef * (0.5 ** Dl), coef * ((1.0 - e) ** Aa) * (e ** Ar)]
        AD_like[c:] = AD_summary_dict[(AD_array[c, 0], AD_array[c, 1])][1]       # `[c:]` as written at zscore.py:78
    return AD_factorial, AD_like, AD_index


def _moments(L, L_keep, A, AD, AD_factorial, AD_like, AD_index, i):
    ctx = _flat_ctx(L, AD)
    return ctx.zmoments_list(i, np.ascontiguousarray(L_keep, dtype=np.int32), np.ascontiguousarray(A, dtype=np.float32),
                             AD_factorial, AD_like, AD_index)


def get_expected_W_l(L, L_keep, A, AD, AD_array, AD_factorial, AD_like, AD_index, t, i):
    """zscore.get_expected_W_l (zscore.py:81-101): ``(W_l_obs float32, W_l_array float32[M_keep])``."""
    w_obs, w_l, _ = _moments(L, L_keep, A, AD, AD_factorial, AD_like, AD_index, i)
    return np.sum(w_obs, dtype=np.float32), w_l


def get_var_W_l(L, L_keep, A, AD, AD_array, AD_factorial, AD_like, AD_index, W_l_array, t, i):
    """zscore.get_var_W_l (zscore.py:103-120): float32 ``[M_keep]`` per-site variances.  The
    kernel recomputes the site's expectation (identical bits to ``W_l_array``) in the same
    launch, so the argument is accepted for signature compatibility only."""
    return _moments(L, L_keep, A, AD, AD_factorial, AD_like, AD_index, i)[2]
