"""ctypes binding of ``libwgsassign_b200.so`` (C ABI in ``include/wgsassign_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There
is no CPU fallback: if the library is missing or no CUDA device is present every operator
raises.
"""
import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# WGS_B200_LIB: load another build of the same library (kernel variants compiled with different -D switches)
LIB_PATH = os.environ.get("WGS_B200_LIB") or os.path.join(_PKG, "libwgsassign_b200.so")
SOURCES = [os.path.join(_PKG, "csrc", "wgs_api.cu"), os.path.join(_PKG, "csrc", "wgs_reader.cpp")]
HEADERS = [os.path.join(_PKG, "csrc", "wgs_kernels.cuh"), os.path.join(_PKG, "csrc", "wgs_zscore.cuh"),
           os.path.join(_ROOT, "include", "wgsassign_b200.h")]

ALLREDUCE_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p)
WGS_F64, WGS_I64 = 0, 1


class ZRow(ctypes.Structure):
    _fields_ = [("z", ctypes.c_float), ("w_obs", ctypes.c_float), ("z_mu", ctypes.c_float),
                ("z_var", ctypes.c_float), ("loci_kept", ctypes.c_int64), ("n_classes", ctypes.c_int32),
                ("em_iters", ctypes.c_int32)]


# every symbol include/wgsassign_b200.h declares: (restype, argtypes)
_vp, _i32, _i64, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
SYMBOLS = {
    "wgs_abi_version": (_i32, []),
    "wgs_device_count": (_i32, []),
    "wgs_device_pci_bus_id": (_i32, [_i32, ctypes.c_char_p, _i32]),
    "wgs_last_error": (ctypes.c_char_p, [_vp]),
    "wgs_create": (_i32, [_i32, ctypes.POINTER(_vp)]),
    "wgs_destroy": (None, [_vp]),
    "wgs_set_option": (_i32, [_vp, ctypes.c_char_p, _i32]),
    "wgs_get_option": (_i32, [_vp, ctypes.c_char_p, _vp]),
    "wgs_set_rank": (_i32, [_vp, _i32, _i32]),
    "wgs_partials_combined": (_i32, [_vp]),
    "wgs_host_alloc": (_vp, [_i64]),
    "wgs_host_free": (None, [_vp]),
    "wgs_set_pops": (_i32, [_vp, _vp, _i32, _i32]),
    "wgs_upload_gl": (_i32, [_vp, _vp, _i64, _i32, _i32]),
    "wgs_nccl_unique_id": (_i32, [_vp]),
    "wgs_nccl_init": (_i32, [_vp, _vp, _i32, _i32]),
    "wgs_upload_gl_async": (_i32, [_vp, _vp, _i64, _i32]),
    "wgs_upload_wait": (_i32, [_vp]),
    "wgs_upload_gl_begin": (_i32, [_vp, _i64, _i32]),
    "wgs_upload_gl_rows": (_i32, [_vp, _vp, _i64, _i64]),
    "wgs_upload_gl_end": (_i32, [_vp, _i64]),
    "wgs_upload_ad": (_i32, [_vp, _vp, _i64, _i32]),
    "wgs_set_shard": (_i32, [_vp, _i64, _i64, ALLREDUCE_FN, _vp]),
    "wgs_synth": (_i32, [_vp, _i64, _i32, ctypes.c_uint64, ctypes.c_float, _i32]),
    "wgs_download": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "wgs_emMAF": (_i32, [_vp, _vp, _i64, _i32, _i32, _f64, _vp, _vp]),
    "wgs_ref_af": (_i32, [_vp, _i32, _f64, _vp, _vp]),
    "wgs_pop_like_partial": (_i32, [_vp, _vp, _i32, _vp]),
    "wgs_loo_partial": (_i32, [_vp, _vp, _i32, _f64, _i32, _i32, _vp, _vp, _vp]),
    "wgs_ref_af_loo": (_i32, [_vp, _i32, _f64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "wgs_fisher_partial": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "wgs_zscore": (_i32, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f64, _vp]),
    "wgs_zscore_classes": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "wgs_zscore_deep_sites": (_i64, [_vp]),
    "wgs_zscore_table": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "wgs_zkeep_one": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "wgs_zmoments_list": (_i32, [_vp, _i32, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "wgs_upload_ad_u8": (_i32, [_vp, _vp, _i64, _i32]),
    "wgs_beagle_stream_open": (_i32, [ctypes.c_char_p, _i32, ctypes.POINTER(_vp)]),
    "wgs_beagle_stream_inds": (_i32, [_vp]),
    "wgs_beagle_stream_sample": (ctypes.c_char_p, [_vp, _i32]),
    "wgs_beagle_stream_keep": (_i32, [_vp, _i64, _i64]),
    "wgs_beagle_stream_names": (_i32, [_vp, _i32]),
    "wgs_beagle_stream_next": (_i64, [_vp, _vp, _i64]),
    "wgs_beagle_stream_rows_seen": (_i64, [_vp]),
    "wgs_beagle_stream_site": (ctypes.c_char_p, [_vp, _i64]),
    "wgs_beagle_stream_sites_joined": (_i64, [_vp, _vp, _i64]),
    "wgs_beagle_stream_estimate_rows": (_i64, [_vp]),
    "wgs_stream_is_bgzf": (_i32, [_vp]),
    "wgs_beagle_stream_open_part": (_i32, [ctypes.c_char_p, _i32, _i32, _i32, ctypes.POINTER(_vp)]),
    "wgs_beagle_stream_stats": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "wgs_beagle_stream_close": (None, [_vp]),
    "wgs_ad_stream_open": (_i32, [ctypes.c_char_p, _i32, ctypes.POINTER(_vp)]),
    "wgs_ad_stream_inds": (_i32, [_vp]),
    "wgs_ad_stream_keep": (_i32, [_vp, _i64, _i64]),
    "wgs_ad_stream_next_u8": (_i64, [_vp, _vp, _i64]),
    "wgs_ad_stream_next_i32": (_i64, [_vp, _vp, _i64]),
    "wgs_ad_stream_rows_seen": (_i64, [_vp]),
    "wgs_ad_stream_estimate_rows": (_i64, [_vp]),
    "wgs_ad_stream_stats": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "wgs_ad_stream_close": (None, [_vp]),
    "wgs_beagle_open": (_i32, [ctypes.c_char_p, _i32, ctypes.POINTER(_vp)]),
    "wgs_beagle_last_error": (ctypes.c_char_p, []),
    "wgs_beagle_sites": (_i64, [_vp]),
    "wgs_beagle_inds": (_i32, [_vp]),
    "wgs_beagle_sample": (ctypes.c_char_p, [_vp, _i32]),
    "wgs_beagle_site": (ctypes.c_char_p, [_vp, _i64]),
    "wgs_beagle_copy": (_i32, [_vp, _vp]),
    "wgs_beagle_close": (None, [_vp]),
    "wgs_debug_stream": (_i32, [_vp, _i32, _vp, _vp]),
    "wgs_debug_seqsum": (_i32, [_vp, _vp, _i64, ctypes.c_float, _vp]),
    "wgs_launch_count": (_i64, [_vp]),
    "wgs_timing_reset": (_i32, [_vp, _i32]),
    "wgs_timing_get": (_i32, [_vp, ctypes.c_char_p, _vp, _vp]),
    "wgs_timing_work": (_i32, [_vp, ctypes.c_char_p, _vp, _vp]),
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-lz", "-split-compile", "0"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS if os.path.exists(p))


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a in-tree (works without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    subprocess.check_call(cmd, cwd=_ROOT)
    return LIB_PATH


_lib = None


def lib():
    """The loaded shared library with typed signatures.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "wgsassign_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.wgs_abi_version() != 1:
            raise RuntimeError("wgsassign_b200: ABI version mismatch")
        _lib = L
    return _lib


class WgsError(RuntimeError):
    pass


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _as(a, dtype, ndim, what):
    """The reference's typed memoryviews raise ValueError on a wrong dtype/layout."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.ndim != ndim or not a.flags.c_contiguous:
        raise ValueError("%s: expected a C-contiguous %s array with %d dimension(s)" % (what, np.dtype(dtype).name, ndim))
    return a


class Context:
    """One GPU's resident state: the repacked GL matrix (+ optional down-sampled matrix and
    allele depths) and the operators over it."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p(0)
        self._cb = None
        L = lib()
        if L.wgs_create(int(device), ctypes.byref(self._h)) != 0:
            raise WgsError(L.wgs_last_error(None).decode())
        self.device = device
        self._pending_L = None
        self.M = 0
        self.N = 0
        self.K = 0

    def close(self):
        if self._h:
            lib().wgs_destroy(self._h)
            self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise WgsError(lib().wgs_last_error(self._h).decode())

    # ---- options ----
    def set_option(self, name, value):
        """Run-time switch of the library (fallback kernels, numerics variants); see kOptionNames in csrc/wgs_api.cu."""
        self._ck(lib().wgs_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name):
        v = ctypes.c_int32(0)
        if lib().wgs_get_option(self._h, name.encode(), ctypes.byref(v)) != 0:
            raise KeyError(name)
        return v.value

    def set_rank(self, rank, world):
        self._ck(lib().wgs_set_rank(self._h, int(rank), int(world)))

    def partials_combined(self):
        """True when the operators' per-individual sums are already summed over the ranks (NCCL on the device)."""
        return bool(lib().wgs_partials_combined(self._h))

    # ---- residency ----
    def set_pops(self, pop_of_ind, K):
        p = np.ascontiguousarray(pop_of_ind, dtype=np.int32)
        self._ck(lib().wgs_set_pops(self._h, _ptr(p), p.shape[0], int(K)))
        self.N, self.K = p.shape[0], int(K)

    def upload_gl(self, L, which=0):
        _as(L, np.float32, 2, "L")
        if L.shape[1] % 2:
            raise ValueError("L must have 2 columns per individual")
        self._ck(lib().wgs_upload_gl(self._h, _ptr(L), L.shape[0], L.shape[1] // 2, int(which)))
        if which == 0:
            self.M, self.N = L.shape[0], L.shape[1] // 2

    def upload_gl_async(self, L):
        """Queue the upload (one strided DMA per population slab) and return at once; the array is
        kept referenced until the next operator (or upload_wait) has consumed it."""
        _as(L, np.float32, 2, "L")
        if L.shape[1] % 2:
            raise ValueError("L must have 2 columns per individual")
        self._pending_L = L
        self._ck(lib().wgs_upload_gl_async(self._h, _ptr(L), L.shape[0], L.shape[1] // 2))
        self.M, self.N = L.shape[0], L.shape[1] // 2

    def upload_gl_begin(self, M_capacity, N):
        """Row-block upload of a matrix that is still being parsed: begin / rows / end (see the header)."""
        self._ck(lib().wgs_upload_gl_begin(self._h, int(M_capacity), int(N)))
        self.N = int(N)

    def upload_gl_rows(self, L, row0, row1):
        """Queue rows [row0, row1) of the float32 [*, 2N] array L (kept referenced until the upload is waited for)."""
        self._pending_L = L
        self._ck(lib().wgs_upload_gl_rows(self._h, ctypes.c_void_p(L.ctypes.data + row0 * L.shape[1] * 4), int(row0), int(row1 - row0)))

    def upload_gl_end(self, M_final):
        self._ck(lib().wgs_upload_gl_end(self._h, int(M_final)))
        self.M = int(M_final)

    def upload_wait(self):
        self._ck(lib().wgs_upload_wait(self._h))
        self._pending_L = None

    def upload_ad(self, AD):
        """AD: int32 [M,2N] (the reference's matrix) or uint8 [M,2N] with 255 = "255 reads or more" (what the
        streaming depth reader produces: a quarter of the bytes)."""
        if isinstance(AD, np.ndarray) and AD.dtype == np.uint8:
            _as(AD, np.uint8, 2, "AD")
            self._ck(lib().wgs_upload_ad_u8(self._h, _ptr(AD), AD.shape[0], AD.shape[1] // 2))
            return
        _as(AD, np.int32, 2, "AD")
        self._ck(lib().wgs_upload_ad(self._h, _ptr(AD), AD.shape[0], AD.shape[1] // 2))

    def set_shard(self, M_total, site_offset, allreduce=None):
        """allreduce(buf: np.ndarray) must sum `buf` across ranks in place."""
        if allreduce is None:
            self._cb = ctypes.cast(None, ALLREDUCE_FN)
        else:
            def _tramp(buf, n, dtype, user, _f=allreduce):
                ct = ctypes.c_double if dtype == WGS_F64 else ctypes.c_int64
                arr = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ct)), shape=(n,))
                _f(arr)
            self._cb = ALLREDUCE_FN(_tramp)
        self._ck(lib().wgs_set_shard(self._h, int(M_total), int(site_offset), self._cb, None))

    def nccl_init(self, uid, rank, world):
        """Collective: give this context an NCCL communicator for the sharded EM stop rule (uid: the 128 bytes
        rank 0 obtained from nccl_unique_id())."""
        buf = ctypes.create_string_buffer(bytes(uid), 128)
        self._ck(lib().wgs_nccl_init(self._h, ctypes.cast(buf, ctypes.c_void_p), int(rank), int(world)))

    def synth(self, M, N, seed=0, depth=2.0, with_ad=False):
        self._ck(lib().wgs_synth(self._h, int(M), int(N), int(seed), float(depth), int(with_ad)))
        self.M, self.N = int(M), int(N)

    def download(self, site0, nsites, want_ad=False):
        L = np.empty((nsites, 2 * self.N), np.float32)
        AD = np.empty((nsites, 2 * self.N), np.int32) if want_ad else None
        self._ck(lib().wgs_download(self._h, int(site0), int(nsites), _ptr(L), _ptr(AD)))
        return (L, AD) if want_ad else L

    # ---- operators ----
    def emMAF(self, L_pop, iters, tole):
        _as(L_pop, np.float32, 2, "L")
        f = np.empty(L_pop.shape[0], np.float32)
        it = ctypes.c_int32(0)
        self._ck(lib().wgs_emMAF(self._h, _ptr(L_pop), L_pop.shape[0], L_pop.shape[1] // 2, int(iters), float(tole),
                                 _ptr(f), ctypes.byref(it)))
        return f, it.value

    def ref_af(self, iters, tole, download=True):
        """Per-population EM + clipping.  With download=False the matrix stays on the device
        only (for a following loo_partial(None, ...)) and None is returned for it."""
        af = np.empty((self.M, self.K), np.float32) if download else None
        its = np.zeros(self.K, np.int32)
        self._ck(lib().wgs_ref_af(self._h, int(iters), float(tole), _ptr(af), _ptr(its)))
        return af, its

    def pop_like_partial(self, af):
        _as(af, np.float32, 2, "af")
        if af.shape[0] != self.M:
            raise ValueError("af has %d rows, the GL matrix %d sites" % (af.shape[0], self.M))
        out = np.empty((self.N, af.shape[1]), np.float64)
        self._ck(lib().wgs_pop_like_partial(self._h, _ptr(af), af.shape[1], _ptr(out)))
        return out

    def loo_partial(self, af, iters, tole, use_ds=False, parts=1):
        """af: float32 [M,K] (mutated in place like glassy.py:89) or None to use the matrix
        the last ref_af left on the device."""
        if af is not None:
            _as(af, np.float32, 2, "af")
            if af.shape != (self.M, self.K):
                raise ValueError("af must be [M,K] = %s" % ((self.M, self.K),))
        ll = np.empty((self.N, self.K), np.float64)
        llp = np.empty((self.N * parts, self.K), np.float64)
        its = np.zeros(self.N, np.int32)
        self._ck(lib().wgs_loo_partial(self._h, _ptr(af), int(iters), float(tole), int(bool(use_ds)), int(parts),
                                       _ptr(ll), _ptr(llp), _ptr(its)))
        return ll, llp, its

    def ref_af_loo(self, iters, tole, use_ds=False, parts=1, want_af_after=False, af_out=None):
        """`--get_reference_af --loo` in one call: (af [M,K], af_iters [K], ll [N,K] f64, ll_parts, loo_iters [N],
        af_after_loo or None) - bit-identical to ref_af() followed by loo_partial(); after upload_gl_async the
        leave-one-out EM overlaps the upload.  af_out: optional preallocated [M,K] float32 (e.g. pinned_empty) to
        receive the allele frequencies."""
        af = np.empty((self.M, self.K), np.float32) if af_out is None else _as(af_out, np.float32, 2, "af_out")
        if af.shape != (self.M, self.K):
            raise ValueError("af_out must be [M,K] = %s" % ((self.M, self.K),))
        af_after = np.empty((self.M, self.K), np.float32) if want_af_after else None
        its = np.zeros(self.K, np.int32)
        ll = np.empty((self.N, self.K), np.float64)
        llp = np.empty((self.N * parts, self.K), np.float64)
        lits = np.zeros(self.N, np.int32)
        self._ck(lib().wgs_ref_af_loo(self._h, int(iters), float(tole), _ptr(af), _ptr(its), _ptr(af_after),
                                      int(bool(use_ds)), int(parts), _ptr(ll), _ptr(llp), _ptr(lits)))
        self._pending_L = None
        return af, its, ll, llp, lits, af_after

    def fisher_partial(self, af):
        _as(af, np.float32, 2, "af")
        if af.shape != (self.M, self.K):
            raise ValueError("af must be [M,K] = %s" % ((self.M, self.K),))
        f_obs = np.empty((self.M, self.K), np.float32)
        ne = np.empty((self.M, self.K), np.float32)
        ind = np.empty(self.N, np.float64)
        self._ck(lib().wgs_fisher_partial(self._h, _ptr(af), _ptr(f_obs), _ptr(ne), _ptr(ind)))
        return f_obs, ne, ind

    def zscore(self, mode, af, n_threshold, single_read, ind_start, ind_end, iters=200, tole=1e-4):
        n = ind_end - ind_start
        rows = (ZRow * max(n, 1))()
        K = 0 if af is None else af.shape[1]
        if af is not None:
            _as(af, np.float32, 2, "af")
            if af.shape != (self.M, self.K):          # the reference raises IndexError on a short AF file; never read past the buffer
                raise ValueError("af must be [M,K] = %s, got %s" % ((self.M, self.K), af.shape))
        self._ck(lib().wgs_zscore(self._h, int(mode), _ptr(af), K, int(n_threshold), int(bool(single_read)),
                                  int(ind_start), int(ind_end), int(iters), float(tole), ctypes.byref(rows)))
        return [rows[i] for i in range(n)]

    def zscore_classes(self, ind, max_rows=4096):
        buf = np.zeros((max_rows, 4), np.int32)
        n = ctypes.c_int32(0)
        self._ck(lib().wgs_zscore_classes(self._h, int(ind), max_rows, _ptr(buf), ctypes.byref(n)))
        return buf[:n.value].copy()

    def zscore_table(self, ind, max_rows=4096):
        buf = np.zeros((max_rows, 7), np.float32)
        n = ctypes.c_int32(0)
        self._ck(lib().wgs_zscore_table(self._h, int(ind), max_rows, _ptr(buf), ctypes.byref(n)))
        return buf[:n.value].copy()

    def zkeep_one(self, ind, ad_array, class_means):
        ad = np.ascontiguousarray(ad_array, dtype=np.int32).reshape(-1, 4)
        cm = np.ascontiguousarray(class_means, dtype=np.float32).reshape(-1, 3)
        out = np.empty(self.M, np.int32)
        n = ctypes.c_int64(0)
        self._ck(lib().wgs_zkeep_one(self._h, int(ind), ad.shape[0], _ptr(ad), _ptr(cm), _ptr(out), self.M, ctypes.byref(n)))
        return out[:n.value].copy()

    def zmoments_list(self, ind, L_keep, A_vec, AD_factorial, AD_like, AD_index):
        keep = _as(L_keep, np.int32, 1, "L_keep")
        a = _as(A_vec, np.float32, 1, "A")
        fac = _as(AD_factorial, np.float32, 2, "AD_factorial")
        like = _as(AD_like, np.float32, 2, "AD_like")
        idx = _as(AD_index, np.int32, 2, "AD_index")
        mk = keep.shape[0]
        w_obs, w_l, var = (np.empty(mk, np.float32) for _ in range(3))
        self._ck(lib().wgs_zmoments_list(self._h, int(ind), _ptr(keep), mk, _ptr(a), fac.shape[0], _ptr(fac), _ptr(like), _ptr(idx),
                                         idx.shape[0], idx.shape[1], _ptr(w_obs), _ptr(w_l), _ptr(var)))
        return w_obs, w_l, var

    def zscore_deep_sites(self):
        return int(lib().wgs_zscore_deep_sites(self._h))

    def debug_seqsum(self, x, carry_in=0.0):
        """Order-exact float32 sum of `x` on the device (the stop rule's primitive); float32 scalar."""
        x = _as(np.ascontiguousarray(x, dtype=np.float32), np.float32, 1, "x")
        out = ctypes.c_float(0)
        self._ck(lib().wgs_debug_seqsum(self._h, _ptr(x), x.shape[0], float(carry_in), ctypes.byref(out)))
        return np.float32(out.value)

    def debug_stream(self, mode):
        ms = ctypes.c_double(0)
        nb = ctypes.c_double(0)
        self._ck(lib().wgs_debug_stream(self._h, int(mode), ctypes.byref(ms), ctypes.byref(nb)))
        return ms.value, nb.value

    # ---- instrumentation ----
    def launch_count(self):
        return int(lib().wgs_launch_count(self._h))

    def timing_reset(self, enable=True):
        self._ck(lib().wgs_timing_reset(self._h, int(enable)))

    def timing_get(self, name):
        ms = ctypes.c_double(0)
        n = ctypes.c_int64(0)
        self._ck(lib().wgs_timing_get(self._h, name.encode(), ctypes.byref(ms), ctypes.byref(n)))
        b = ctypes.c_double(0)
        u = ctypes.c_double(0)
        self._ck(lib().wgs_timing_work(self._h, name.encode(), ctypes.byref(b), ctypes.byref(u)))
        return dict(ms=ms.value, launches=n.value, bytes=b.value, units=u.value)


def nccl_unique_id():
    buf = ctypes.create_string_buffer(128)
    if lib().wgs_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p)) != 0:
        raise WgsError(lib().wgs_last_error(None).decode())
    return buf.raw


def pinned_empty(shape, dtype):
    """A numpy array in pinned host memory (full-rate, asynchronous H2D).  The pages are released
    (cudaFreeHost) when the last view of the buffer is garbage-collected."""
    import weakref
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = lib().wgs_host_alloc(max(nbytes, 1))
    if not p:
        raise WgsError("cudaHostAlloc failed")
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(p)
    # every array made from `buf` keeps it alive through .base; the finalizer runs when `buf` itself dies
    weakref.finalize(buf, lib().wgs_host_free, p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
