"""TEST INFRASTRUCTURE - CPU oracle for the WGSassign genotype-likelihood hot path.

Restates the reference's Python drivers (emMAF.py, glassy.py, fisher.py, zscore.py,
mixture.py and the per-population / per-individual loops of WGSassign.py) on top of one
of two kernel backends:

* ``port`` - ``oracle/liboracle.so``, the plain-C restatement in ``oracle/oracle.c``;
* ``ref``  - the reference's OWN compiled Cython kernels in ``oracle/_ref`` (built from
  /root/reference by ``oracle/build_ref.sh``; binaries only, git-ignored).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product (``wgsassign_b200``) never does.

Parity status: PINNED - ``tests/test_oracle.py`` checks both backends bit-for-bit against
each other, against the unmodified reference drivers when /root/reference is present, and
against the golden fixtures in ``tests/golden`` produced by the reference CLI.

All ``file:line`` citations are relative to /root/reference/WGSassign/.
"""
import ctypes
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def _fp(a):
    return a.ctypes.data_as(_F)


def _ip(a):
    return a.ctypes.data_as(_I)


def _chk(a, dtype, ndim):
    # The reference's typed memoryviews reject anything else (SURVEY 8b "Types").
    if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.ndim == ndim and a.flags.c_contiguous):
        raise ValueError("ndarray is not C-contiguous %s[%dd]" % (np.dtype(dtype).name, ndim))


class PortKernels:
    """ctypes binding of oracle/liboracle.so with the reference's cpdef signatures."""

    name = "port"

    def __init__(self, path=None):
        path = path or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing - run `make -C oracle liboracle.so`")
        self.lib = ctypes.CDLL(path)
        self.lib.orc_rmse1d.restype = ctypes.c_double
        self.lib.orc_emMAF.restype = ctypes.c_int

    # emMAF_cy.pyx:10 / :26
    def emMAF_update(self, L, f, t):
        _chk(L, np.float32, 2); _chk(f, np.float32, 1)
        self.lib.orc_em_update(_fp(L), L.shape[0], L.shape[1] // 2, _fp(f), int(t))

    def rmse1d(self, v1, v2):
        _chk(v1, np.float32, 1); _chk(v2, np.float32, 1)
        return float(self.lib.orc_rmse1d(_fp(v1), _fp(v2), v1.shape[0]))

    def emMAF_full(self, L, iters, tole, t):
        _chk(L, np.float32, 2)
        m = L.shape[0]
        f = np.empty(m, np.float32); prev = np.empty(m, np.float32)
        it = self.lib.orc_emMAF(_fp(L), m, L.shape[1] // 2, int(iters), ctypes.c_double(tole),
                                _fp(f), _fp(prev), int(t))
        return f, int(it)

    # glassy_cy.pyx:12
    def loglike(self, L, A, vec, t, i, k):
        _chk(L, np.float32, 2); _chk(A, np.float32, 2); _chk(vec, np.float32, 1)
        self.lib.orc_loglike(_fp(L), L.shape[0], L.shape[1], _fp(A), A.shape[1], int(i), int(k),
                             _fp(vec), int(t))

    # fisher_cy.pyx:12 / :32 / :41 / :58
    def fisher_obs(self, L, A, t, i, n, f_pop):
        _chk(L, np.float32, 2); _chk(A, np.float32, 2); _chk(f_pop, np.float32, 1)
        self.lib.orc_fisher_obs(_fp(L), A.shape[0], int(n), _fp(A), A.shape[1], int(i), _fp(f_pop), int(t))

    def ne_obs(self, f_pop, A, t, i, n, ne_pop):
        self.lib.orc_ne_obs(_fp(f_pop), A.shape[0], _fp(A), A.shape[1], int(i), _fp(ne_pop), int(t))

    def fisher_obs_ind(self, L, A, t, i, pop_i, f_ind):
        _chk(L, np.float32, 2); _chk(A, np.float32, 2)
        self.lib.orc_fisher_obs_ind(_fp(L), A.shape[0], L.shape[1], _fp(A), A.shape[1], int(i),
                                    int(pop_i), _fp(f_ind), int(t))

    def ne_obs_ind(self, f_ind, A, t, pop_i, ne_ind):
        self.lib.orc_ne_obs(_fp(f_ind), A.shape[0], _fp(A), A.shape[1], int(pop_i), _fp(ne_ind), int(t))

    # zscore_cy.pyx:10 / :37
    def expected_W_l(self, L, L_keep, A, AD, AD_array, AD_factorial, AD_like, AD_index, t, i, W_obs, W_l):
        _chk(L, np.float32, 2); _chk(L_keep, np.int32, 1); _chk(A, np.float32, 1); _chk(AD, np.int32, 2)
        _chk(AD_factorial, np.float32, 2); _chk(AD_like, np.float32, 2); _chk(AD_index, np.int32, 2)
        self.lib.orc_expected_W_l(_fp(L), L.shape[1], _ip(L_keep), L_keep.shape[0], _fp(A), _ip(AD),
                                  _fp(AD_factorial), _fp(AD_like), _ip(AD_index), AD_index.shape[1],
                                  int(i), _fp(W_obs), _fp(W_l), int(t))

    def variance_W_l(self, L, L_keep, A, AD, AD_array, AD_factorial, AD_like, AD_index, t, i, var, W_l):
        _chk(L, np.float32, 2); _chk(L_keep, np.int32, 1); _chk(A, np.float32, 1); _chk(AD, np.int32, 2)
        self.lib.orc_variance_W_l(_fp(L), L.shape[1], _ip(L_keep), L_keep.shape[0], _fp(A), _ip(AD),
                                  _fp(AD_factorial), _fp(AD_like), _ip(AD_index), AD_index.shape[1],
                                  int(i), _fp(var), _fp(W_l), int(t))


class RefKernels:
    """The reference's own compiled Cython kernels (oracle/_ref/WGSassign/*.so)."""

    name = "ref"

    def __init__(self):
        root = os.path.join(_HERE, "_ref")
        if not os.path.isdir(os.path.join(root, "WGSassign")):
            raise RuntimeError("oracle/_ref missing - run oracle/build_ref.sh where /root/reference exists")
        if root not in sys.path:
            sys.path.insert(0, root)
        from WGSassign import emMAF_cy, fisher_cy, glassy_cy, zscore_cy  # noqa: E402
        self.emMAF_update = emMAF_cy.emMAF_update
        self.rmse1d = emMAF_cy.rmse1d
        self.loglike = glassy_cy.loglike
        self.fisher_obs = fisher_cy.fisher_obs
        self.ne_obs = fisher_cy.ne_obs
        self.fisher_obs_ind = fisher_cy.fisher_obs_ind
        self.ne_obs_ind = fisher_cy.ne_obs_ind
        self.expected_W_l = zscore_cy.expected_W_l
        self.variance_W_l = zscore_cy.variance_W_l

    def emMAF_full(self, L, iters, tole, t):
        # emMAF.py:15-27
        m = L.shape[0]
        f = np.full(m, 0.25, np.float32)
        prev = f.copy()
        for it in range(iters):
            self.emMAF_update(L, f, t)
            if self.rmse1d(f, prev) < tole:
                return f, it + 1
            prev = f.copy()
        return f, 0


def have_ref():
    return os.path.isdir(os.path.join(_HERE, "_ref", "WGSassign"))


_cache = {}


def kernels(backend="port"):
    if backend == "auto":
        backend = "ref" if have_ref() else "port"
    if backend not in _cache:
        _cache[backend] = PortKernels() if backend == "port" else RefKernels()
    return _cache[backend]


# --------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------
def pop_cols(IDs, pop, exclude=None):
    """Column indices (g0,g1 interleaved, ascending) of a population's individuals
    (WGSassign.py:227-232, glassy.py:69-76)."""
    idx = np.flatnonzero(IDs[:, 1] == pop)
    if exclude is not None:
        idx = idx[idx != exclude]
    cols = np.empty(2 * idx.size, np.int64)
    cols[0::2] = 2 * idx
    cols[1::2] = 2 * idx + 1
    return cols


def clip_af(f, n_pop):
    """WGSassign.py:236-240 / glassy.py:80-85: clamp to [1/(2(n+1)), 1-1/(2(n+1))]."""
    lo = 1 / (2 * (n_pop + 1))
    hi = 1 - lo
    f[f < lo] = lo
    f[f > hi] = hi
    return f


def emMAF(L, iters, tole, t, kern=None):
    """emMAF.py:15-27.  Returns (f, converged_iteration or 0)."""
    kern = kern or kernels()
    return kern.emMAF_full(np.ascontiguousarray(L), iters, tole, t)


def reference_af(L, IDs, iters, tole, t, kern=None):
    """WGSassign.py:213-242.  Returns (af [M,K] float32, pops, iterations per pop)."""
    kern = kern or kernels()
    pops = np.unique(IDs[:, 1])
    af = np.empty((L.shape[0], len(pops)), np.float32)
    its = []
    for k, p in enumerate(pops):
        L_pop = np.ascontiguousarray(L[:, pop_cols(IDs, p)])
        f, it = emMAF(L_pop, iters, tole, t, kern)
        af[:, k] = clip_af(f, L_pop.shape[1] // 2)
        its.append(it)
    return af, pops, its


def assignLL(L, af, t, kern=None):
    """glassy.py:18-44: float64 pairwise sum of the per-site float32 vector, stored float32."""
    kern = kern or kernels()
    m, n, k = L.shape[0], L.shape[1] // 2, af.shape[1]
    out = np.zeros((n, k), np.float32)
    for i in range(n):
        for j in range(k):
            vec = np.zeros(m, np.float32)
            kern.loglike(L, af, vec, t, i, j)
            out[i, j] = np.sum(vec, dtype=float)
    return out


def partition_loglikes(vec, parts):
    """utils.py:129-151: modulo partition sums, unbuffered float32 np.add.at."""
    labels = np.arange(vec.shape[0]) % parts
    out = np.zeros(parts, np.float32)
    np.add.at(out, labels, vec)
    return out


def loo(L, af, IDs, t, iters, tole, downsampled_L=None, num_partitions=1, kern=None):
    """glassy.py:47-112.  Mutates `af` in place exactly like the reference (:89).
    Returns (logl_mat, logl_parts_mat, EM iterations per individual)."""
    kern = kern or kernels()
    m, n, k = L.shape[0], L.shape[1] // 2, af.shape[1]
    out = np.zeros((n, k), np.float32)
    parts = np.zeros((n * num_partitions, k), np.float32)
    pops = np.unique(IDs[:, 1])
    src = downsampled_L if downsampled_L is not None else L
    its = []
    for i in range(n):
        L_pop = np.ascontiguousarray(L[:, pop_cols(IDs, IDs[i, 1], exclude=i)])
        f, it = emMAF(L_pop, iters, tole, t, kern)
        its.append(it)
        col = int(np.flatnonzero(pops == IDs[i, 1])[0])
        af[:, col] = clip_af(f, L_pop.shape[1] // 2)
        for j in range(k):
            vec = np.zeros(m, np.float32)
            kern.loglike(src, af, vec, t, i, j)
            out[i, j] = np.sum(vec, dtype=float)
            parts[i * num_partitions:(i + 1) * num_partitions, j] = partition_loglikes(vec, num_partitions)
    return out, parts, its


def fisher_obs(L, af, IDs, t, kern=None):
    """fisher.py:11-43."""
    kern = kern or kernels()
    m = L.shape[0]
    pops = np.unique(IDs[:, 1])
    f_obs = np.empty((m, len(pops)), np.float32)
    ne = np.empty((m, len(pops)), np.float32)
    for i, p in enumerate(pops):
        L_pop = np.ascontiguousarray(L[:, pop_cols(IDs, p)])
        n = L_pop.shape[1] // 2
        f_pop = np.zeros(m, np.float32)
        kern.fisher_obs(L_pop, af, t, i, n, f_pop)
        f_obs[:, i] = f_pop
        ne_pop = np.zeros(m, np.float32)
        kern.ne_obs(f_pop, af, t, i, n, ne_pop)
        ne[:, i] = ne_pop
    return f_obs, ne


def fisher_obs_ind(L, af, IDs, t, kern=None):
    """fisher.py:45-59: per-individual mean over sites (float32 np.mean)."""
    kern = kern or kernels()
    m, n = L.shape[0], L.shape[1] // 2
    pops = np.unique(IDs[:, 1])
    out = np.zeros(n, np.float32)
    for i in range(n):
        f_ind = np.zeros(m, np.float32)
        pop_i = int(np.flatnonzero(pops == IDs[i, 1])[0])
        kern.fisher_obs_ind(L, af, t, i, pop_i, f_ind)
        ne_ind = np.zeros(m, np.float32)
        kern.ne_obs_ind(f_ind, af, t, pop_i, ne_ind)
        out[i] = out[i] + np.mean(ne_ind)
    return out


# ---- z-score ---------------------------------------------------------------------------
def AD_summary(L, AD, i, n_threshold, single_read_threshold):
    """zscore.py:11-41, vectorised but value-identical: classes in first-occurrence order,
    class mean GL = float32 np.mean(axis=0) over that class's rows in site order."""
    ar = AD[:, 2 * i].astype(np.int64)
    aa = AD[:, 2 * i + 1].astype(np.int64)
    g0 = L[:, 2 * i]
    g1 = L[:, 2 * i + 1]
    g2 = (np.float32(1) - g0) - g1                       # :17 float32 scalar arithmetic
    width = int(aa.max()) + 1 if aa.size else 1
    code = ar * width + aa
    uniq, first, inv = np.unique(code, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    summary = {}
    rows_out = []
    gl = np.stack([g0, g1, g2], axis=1).astype(np.float32)
    for u in order:
        rows = np.flatnonzero(inv == u)
        key = (int(ar[rows[0]]), int(aa[rows[0]]))
        mean = np.mean(np.ascontiguousarray(gl[rows]), axis=0)   # :22
        summary[key] = [int(rows.size), mean]
        rows_out.append([key[0], key[1], key[0] + key[1], int(rows.size)])
    arr = np.array(rows_out, np.int32).reshape(-1, 4)
    if single_read_threshold:
        filt = arr[arr[:, 2] == 1]                                   # :31
    else:
        filt = arr[(arr[:, 3] > n_threshold) & (arr[:, 2] != 0)]     # :33
    assert filt.shape[0] != 0, "No loci were kept! Too stringent filtering?"
    assert filt.shape[0] != 1, "Not enough loci were kept! Too stringent filtering?"
    dl, dl_counts = np.unique(filt[:, 0] + filt[:, 1], return_counts=True)
    dl_keep = dl[dl < dl_counts]                                     # :38
    AD_array = filt[np.isin(filt[:, 2], dl_keep)]                    # :39
    return summary, AD_array


def get_L_keep(L, AD, summary, AD_array, i):
    """zscore.py:43-61, vectorised but value-identical."""
    m = AD.shape[0]
    ar = AD[:, 2 * i]
    aa = AD[:, 2 * i + 1]
    g0 = L[:, 2 * i]
    g1 = L[:, 2 * i + 1]
    g2 = (np.float32(1) - g0) - g1
    gl = np.stack([g0, g1, g2], axis=1)
    keep = np.zeros(m, bool)
    for r in range(AD_array.shape[0]):
        key = (int(AD_array[r, 0]), int(AD_array[r, 1]))
        mean = summary[key][1]
        max_id = int(np.argwhere(mean == np.max(mean))[0][0])        # :53
        rows = np.flatnonzero((ar == key[0]) & (aa == key[1]))
        bad = np.abs(mean[max_id] - gl[rows, max_id]) > 0.01         # :55 (float32 compare)
        keep[rows] = ~bad
    L_keep = np.flatnonzero(keep).astype(np.int32)
    return L_keep, int(L_keep.shape[0])


def get_factorials(AD_array, summary, e):
    """zscore.py:63-79."""
    C = AD_array.shape[0]
    fac = np.zeros((C, 3), np.float32)
    like = np.zeros((C, 3), np.float32)
    index = np.zeros((int(np.max(AD_array[:, 0])) + 1, int(np.max(AD_array[:, 1])) + 1), np.int32)
    for c in range(C):
        Ar, Aa = int(AD_array[c, 0]), int(AD_array[c, 1])
        index[Ar, Aa] = c                                            # :71
        Dl = Aa + Ar
        comb = math.factorial(Dl) / (math.factorial(Aa) * math.factorial(Ar))
        fac[c, :] = [comb * ((1.0 - e) ** Ar) * (e ** Aa), comb * (0.5 ** Dl),
                     comb * ((1.0 - e) ** Aa) * (e ** Ar)]           # :74-77
        like[c:] = summary[(Ar, Aa)][1]                              # :78 (slice-to-end quirk)
    return fac, like, index


def zscore_individual(L, AD, i, af_vec_fn, n_threshold, single_read, t, kern=None, e=0.01):
    """One pass of the per-individual z-score body shared by WGSassign.py:346-381 and
    :425-443.  `af_vec_fn(L_keep)` returns the float32 AF vector over kept sites."""
    kern = kern or kernels()
    summary, AD_array = AD_summary(L, AD, i, n_threshold, single_read)
    L_keep, kept = get_L_keep(L, AD, summary, AD_array, i)
    fac, like, index = get_factorials(AD_array, summary, e)
    af_vec = np.ascontiguousarray(af_vec_fn(L_keep), dtype=np.float32)
    W_obs_arr = np.zeros(kept, np.float32)
    W_l = np.zeros(kept, np.float32)
    kern.expected_W_l(L, L_keep, af_vec, AD, AD_array, fac, like, index, t, i, W_obs_arr, W_l)
    W_obs = np.sum(W_obs_arr, dtype=np.float32)                      # zscore.py:100
    var = np.zeros(kept, np.float32)
    kern.variance_W_l(L, L_keep, af_vec, AD, AD_array, fac, like, index, t, i, var, W_l)
    z_mu = np.sum(W_l)
    z_var = np.sum(var)
    with np.errstate(invalid="ignore", divide="ignore"):
        z = (W_obs - z_mu) / np.sqrt(z_var)                          # WGSassign.py:371
    return dict(z=np.float32(z), z_mu=np.float32(z_mu), z_var=np.float32(z_var),
                w_obs=np.float32(W_obs), loci_kept=kept, AD_array=AD_array, L_keep=L_keep,
                af=af_vec, summary=summary)


def zscore_assignment(L, AD, A, IDs, pops, n_threshold=0, single_read=False, ind_start=0,
                      ind_end=None, t=1, kern=None):
    """WGSassign.py:386-446 (--get_assignment_z_score)."""
    n = L.shape[1] // 2
    ind_end = n if ind_end is None else ind_end
    res = []
    for i in range(ind_start, ind_end):
        k = int(np.argwhere(pops == IDs[i, 1])[0][0])
        res.append(zscore_individual(
            L, AD, i, lambda keep: A[keep, :][:, k].reshape(-1), n_threshold, single_read, t, kern))
    return res


def zscore_reference(L, AD, IDs, iters, tole, n_threshold=0, single_read=False, ind_start=0,
                     ind_end=None, t=1, kern=None):
    """WGSassign.py:311-384 (--get_reference_z_score): LOO EM on the kept sites only."""
    kern = kern or kernels()
    n = L.shape[1] // 2
    ind_end = n if ind_end is None else ind_end
    res = []
    for i in range(ind_start, ind_end):
        cols = pop_cols(IDs, IDs[i, 1], exclude=i)
        it_box = []

        def af_fn(keep, cols=cols, it_box=it_box):
            L_pop = np.ascontiguousarray(L[keep, :][:, cols])        # :358
            f, it = emMAF(L_pop, iters, tole, t, kern)
            it_box.append(it)
            return clip_af(f, L_pop.shape[1] // 2)

        r = zscore_individual(L, AD, i, af_fn, n_threshold, single_read, t, kern)
        r["em_iter"] = it_box[0]
        res.append(r)
    return res


def em_mix(L_mat, L_mat_index, iters):
    """mixture.py:10-39."""
    n_source = L_mat.shape[1]
    harvest = np.unique(L_mat_index[:, 1])
    out = np.empty((len(harvest), n_source), np.float32)
    for h, name in enumerate(harvest):
        Lp = np.ascontiguousarray(L_mat[np.flatnonzero(L_mat_index[:, 1] == name), :])
        n_ind = Lp.shape[0]
        pi_mat = np.diag(np.full(n_source, 1)) / n_source
        pi_vec = None
        for _ in range(iters):
            w = np.matmul(np.exp(Lp), pi_mat)
            w = w / w.sum(axis=1, keepdims=True)
            pi_vec = w.sum(axis=0, keepdims=True) / n_ind
            pi_mat = np.diag(pi_vec.reshape(-1))
        out[h, :] = pi_vec
    return np.hstack((harvest.reshape((len(harvest), 1)), out))
