"""Resident-matrix cache behind the drop-in entry points.

The reference's drivers take the GL matrix as a NumPy array on every call.  Re-uploading
it for each of `--get_reference_af`, `--ne_obs` and `--loo` would triple the PCIe traffic,
so the entry points ask this module for a :class:`~wgsassign_b200._lib.Context` that
already holds the matrix; it is re-uploaded only when the array or the population
assignment changes.

How "the array changed" is decided (`_sig`): address, shape, dtype and a checksum of the
contents - over EVERY byte for arrays up to `FULL_CHECK_BYTES` (64 MB: a few milliseconds)
and always when `strict(True)` is set; above that a 65,536-element strided sample plus the
first and last 4 KB, because a full pass over a 4 GB matrix costs several uploads.  A caller
that edits a large matrix in place between two entry-point calls must therefore call
`invalidate()` (or run with `strict(True)`); the CLI never does that.
"""
import os

import numpy as np

from . import _lib
from . import dist

_state = {"ctx": None, "key": None, "ds_key": None, "ad_key": None}
_options = {}                     # library switches applied to every context this module creates (set_option)
FULL_CHECK_BYTES = 64 << 20
_strict = [bool(os.environ.get("WGS_SESSION_STRICT"))]


def strict(on=True):
    """Full-content checksum for every array, whatever its size (debug mode)."""
    _strict[0] = bool(on)


def invalidate():
    """Forget what is resident: the next entry-point call re-uploads its arrays.  Call this after
    modifying a matrix in place."""
    _state.update(key=None, ds_key=None, ad_key=None)
    from . import fisher
    fisher._last.update(key=None, ne_ind=None)


def set_option(name, value):
    """Library switch (`Context.set_option`) for the cached context and every later one."""
    _options[name] = int(value)
    if _state["ctx"] is not None:
        _state["ctx"].set_option(name, value)


def device_index():
    return int(os.environ.get("LOCAL_RANK", os.environ.get("WGS_DEVICE", "0")))


def _sig(a):
    """Identity of an array's contents: address, shape, dtype and a checksum (see the module docstring)."""
    a = np.asarray(a)
    head = (a.ctypes.data, a.shape, a.dtype.str)
    if not a.flags.c_contiguous or a.nbytes == 0:
        return head + (float(np.sum(a, dtype=np.float64)),)
    raw = a.reshape(-1).view(np.uint8)
    if _strict[0] or a.nbytes <= FULL_CHECK_BYTES:
        n8 = raw.shape[0] // 8 * 8
        words = raw[:n8].view(np.uint64)
        # order-sensitive: a weighted sum (index mod 2^16 + 1) catches swapped elements as well as changed ones
        w = (np.arange(words.shape[0], dtype=np.uint64) & np.uint64(0xFFFF)) + np.uint64(1)
        return head + (int(np.sum(words * w, dtype=np.uint64)), int(np.sum(raw[n8:], dtype=np.uint64)))
    flat = a.reshape(-1)
    step = max(1, flat.shape[0] // 65536)
    return head + (float(np.sum(flat[::step], dtype=np.float64)), int(np.sum(raw[:4096], dtype=np.uint64)),
                   int(np.sum(raw[-4096:], dtype=np.uint64)))


def pops_from_ids(IDs):
    """(pop_of_ind int32 [N], pops) with populations in np.unique order (WGSassign.py:213)."""
    pops, inv = np.unique(IDs[:, 1], return_inverse=True)
    return inv.astype(np.int32), pops


def context(L, pop_of_ind=None, K=0, async_upload=False):
    """Context holding `L` repacked for the given population assignment (None = flat).
    async_upload: queue the upload slab by slab and return at once (the next operator waits for
    what it needs; `Context.ref_af_loo` overlaps its leave-one-out EM with the transfer)."""
    pkey = None if pop_of_ind is None else (int(K), np.asarray(pop_of_ind, np.int32).tobytes())
    key = (_sig(L), pkey)
    ctx = _state["ctx"]
    if ctx is not None and _state["key"] == key:
        return ctx
    if ctx is None:
        ctx = _lib.Context(device_index())
        for name, value in _options.items():
            ctx.set_option(name, value)
        _state["ctx"] = ctx
    n = L.shape[1] // 2
    _state.update(key=None, ds_key=None, ad_key=None)      # nothing is trusted until the upload has succeeded
    if pop_of_ind is None:
        ctx.set_pops(np.zeros(n, np.int32), 0)
    else:
        if len(pop_of_ind) != n:
            raise ValueError("Number of individuals in beagle and reference ID file do not match!")
        ctx.set_pops(pop_of_ind, K)
    if async_upload and pop_of_ind is not None:
        ctx.upload_gl_async(L)
    else:
        ctx.upload_gl(L, 0)
    dist.attach(ctx)
    _state["key"] = key
    _state["ds_key"] = None
    _state["ad_key"] = None
    return ctx


def stream_context(beagle, pop_of_ind=None, K=0, threads=0, rows=None, part=None):
    """Parse a Beagle file straight into a resident context: every finished row block is queued for upload while the
    next one is parsed (`readBeagle(on_block=...)` + `Context.upload_gl_begin/rows/end`).  Returns
    (ctx, L, sample_names, site_names); the context is registered for L, so the entry points that are later called
    with this array find it resident.  pop_of_ind must already be known (the device layout is population-sorted).
    part=(p, P): this process reads the p-th byte range of a BGZF file (reader.readBeagle); the shard geometry is then
    known only when every rank has counted its rows, so the CALLER declares it (dist.row_counts_to_ranges, dist.enable)
    and attaches the context (dist.attach) afterwards."""
    from . import reader
    ctx = _state["ctx"]
    if ctx is None:
        ctx = _lib.Context(device_index())
        for name, value in _options.items():
            ctx.set_option(name, value)
        _state["ctx"] = ctx
    _state.update(key=None, ds_key=None, ad_key=None)
    state = {"begun": False, "cap": 0}

    def on_block(arr, r0, r1):
        n = arr.shape[1] // 2
        if not state["begun"] or arr.shape[0] != state["cap"]:      # first block, or the reader grew its buffer: start over
            if pop_of_ind is None:
                ctx.set_pops(np.zeros(n, np.int32), 0)
            else:
                if len(pop_of_ind) != n:
                    raise ValueError("Number of individuals in beagle and reference ID file do not match!")
                ctx.set_pops(pop_of_ind, K)
            if part is None:
                dist.attach(ctx)
            ctx.upload_gl_begin(arr.shape[0], n)
            state.update(begun=True, cap=arr.shape[0])
            if r0 > 0:
                ctx.upload_gl_rows(arr, 0, r0)
        ctx.upload_gl_rows(arr, r0, r1)
    L, samples, sites = reader.readBeagle(beagle, threads, rows=rows, on_block=on_block, part=part)
    if state["begun"]:
        ctx.upload_gl_end(L.shape[0])
        ctx._pending_L = L
        pkey = None if pop_of_ind is None else (int(K), np.asarray(pop_of_ind, np.int32).tobytes())
        _state["key"] = (_sig(L), pkey)
    return ctx, L, samples, sites


def with_downsampled(ctx, L_ds):
    key = _sig(L_ds)
    if _state["ds_key"] != key:
        _state["ds_key"] = None
        ctx.upload_gl(L_ds, 1)
        _state["ds_key"] = key
    return ctx


def with_ad(ctx, AD):
    key = _sig(AD)
    if _state["ad_key"] != key:
        _state["ad_key"] = None                            # a failed upload must not leave a stale key behind
        # uint8 (255 = "255 reads or more", the streaming reader's form) goes up as it is; anything else as the reference's int32
        ctx.upload_ad(np.ascontiguousarray(AD) if AD.dtype == np.uint8 else np.ascontiguousarray(AD, dtype=np.int32))
        _state["ad_key"] = key
    return ctx


def reset():
    ctx = _state["ctx"]
    if ctx is not None:
        ctx.close()
    _state.update(ctx=None, key=None, ds_key=None, ad_key=None)
