"""Drop-in for the reference's ``reader_cy.readBeagle`` (reader_cy.pyx:16-77) and a fast reader for the
``--ind_ad_file`` matrix (the reference's ``np.loadtxt``, WGSassign.py:320, :399), backed by the streaming C++
parsers in ``csrc/wgs_reader.cpp``: a background thread inflates while a persistent thread pool converts rows
straight into the output array (pinned host memory when a CUDA device is present, so the matrix can be uploaded
asynchronously); no ``gunzip`` subprocess, no shell interpolation of the path.

Site-sharded runs (``rows=(lo, hi)``) convert only their own row range; the other rows are counted and named.
"""
import ctypes
import time

import numpy as np

from . import _lib

last_stats = {}          # throughput of the last readBeagle / readAD call (bench.py and the CLI's --verbose report it)


def _alloc(shape, dtype):
    """Pinned host memory when a CUDA device is there (full-rate, asynchronous uploads); plain otherwise."""
    try:
        return _lib.pinned_empty(shape, dtype)
    except Exception:
        return np.empty(shape, dtype)


def _stats(L, h, fn, wall, rows, kind):
    a, b = ctypes.c_double(0), ctypes.c_double(0)
    c, u = ctypes.c_int64(0), ctypes.c_int64(0)
    fn(h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(u))
    last_stats[kind] = dict(wall_s=wall, inflate_s=a.value, parse_s=b.value, compressed_bytes=c.value, uncompressed_bytes=u.value,
                            rows=rows, compressed_mb_per_s=c.value / 1e6 / max(wall, 1e-9),
                            uncompressed_mb_per_s=u.value / 1e6 / max(wall, 1e-9), bgzf=bool(L.wgs_stream_is_bgzf(h)))


def _site_names(L, h):
    """Names of all rows the stream has seen, through ONE call (a call per name costs more than parsing the row)."""
    need = L.wgs_beagle_stream_sites_joined(h, None, 0)
    if need <= 0:
        return []
    buf = ctypes.create_string_buffer(need)
    L.wgs_beagle_stream_sites_joined(h, buf, need)
    return buf.raw[:need - 1].decode().split("\n")


def count_rows(beagle, threads=0):
    """(M, sample_names, site_names) without converting a single number (one inflate pass): what a site-sharded
    rank needs before it can tell which rows are its own."""
    L = _lib.lib()
    h = ctypes.c_void_p(0)
    if L.wgs_beagle_stream_open(str(beagle).encode(), int(threads), ctypes.byref(h)) != 0:
        raise IOError(L.wgs_beagle_last_error().decode())
    try:
        n = L.wgs_beagle_stream_inds(h)
        L.wgs_beagle_stream_keep(h, 0, 0)
        if L.wgs_beagle_stream_next(h, None, 1 << 62) < 0:
            raise IOError(L.wgs_beagle_last_error().decode())
        m = L.wgs_beagle_stream_rows_seen(h)
        samples = [L.wgs_beagle_stream_sample(h, i).decode() for i in range(n)]
        sites = _site_names(L, h)
    finally:
        L.wgs_beagle_stream_close(h)
    return m, samples, sites


def is_bgzf(path):
    """True for a BGZF (bgzip) file - the container ANGSD writes: independent members that can be inflated in parallel
    and entered in the middle (byte-range parts); False for a plain gzip stream or anything else."""
    try:
        with open(path, "rb") as fh:
            h = fh.read(18)
    except OSError:
        return False
    return len(h) >= 18 and h[:3] == b"\x1f\x8b\x08" and bool(h[3] & 4) and h[12:14] == b"BC"


def readBeagle(beagle, threads=0, rows=None, on_block=None, part=None):
    """Returns ``(L float32 [M, 2N], sample_names list[str], site_names list[str])``.

    rows=(lo, hi): convert only that row range (L is then [hi-lo, 2N]; the names are still those of ALL sites).
    on_block(L, r0, r1): called after rows [r0, r1) of L have been written - the CLI queues their upload from it, so
    the transfer of one block overlaps the parsing of the next.
    part=(p, P) (BGZF files only): read the p-th of P byte ranges of the file - the rows that start inside it; L and the
    site names are then those rows only, and P processes reading the P parts inflate the file once between them
    (with ``rows`` every process inflates all of it, twice: once to count).  The parts in order are the file."""
    L = _lib.lib()
    h = ctypes.c_void_p(0)
    t0 = time.perf_counter()
    if part is not None:
        if rows is not None:
            raise ValueError("rows and part are alternatives")
        rc = L.wgs_beagle_stream_open_part(str(beagle).encode(), int(threads), int(part[0]), int(part[1]), ctypes.byref(h))
    else:
        rc = L.wgs_beagle_stream_open(str(beagle).encode(), int(threads), ctypes.byref(h))
    if rc != 0:
        raise IOError(L.wgs_beagle_last_error().decode())
    try:
        n = L.wgs_beagle_stream_inds(h)
        if rows is not None:
            lo, hi = int(rows[0]), int(rows[1])
            L.wgs_beagle_stream_keep(h, lo, hi)
            cap = max(hi - lo, 0)
        else:
            est = L.wgs_beagle_stream_estimate_rows(h)
            cap = int(est * 1.03) + 1024 if est > 0 else 1 << 16
        out = _alloc((max(cap, 1), 2 * n), np.float32)
        block = max(1, (64 << 20) // max(8 * n, 1))                # ~64 MB of output per call: the upload granularity
        m = 0
        while True:
            if m == cap and rows is None:                          # the estimate was short (it is refined as the file is read): grow
                cap = int(max(cap * 1.5, L.wgs_beagle_stream_estimate_rows(h) * 1.03)) + 1024
                bigger = _alloc((cap, 2 * n), np.float32)
                bigger[:m] = out[:m]
                out = bigger
            # with a row range the stream itself stops converting at `hi` and then runs to the end of the file for the
            # names and the row count (the last call returns 0 without writing)
            want = min(block, cap - m) if rows is None else max(1, min(block, cap - m))
            got = L.wgs_beagle_stream_next(h, ctypes.c_void_p(out.ctypes.data + m * 2 * n * 4), want)
            if got < 0:
                raise IOError(L.wgs_beagle_last_error().decode())
            if got == 0:
                break
            if on_block is not None:
                on_block(out, m, m + got)
            m += got
        total = L.wgs_beagle_stream_rows_seen(h)
        samples = [L.wgs_beagle_stream_sample(h, i).decode() for i in range(n)]
        sites = _site_names(L, h)
        _stats(L, h, L.wgs_beagle_stream_stats, time.perf_counter() - t0, total, "beagle")
    finally:
        L.wgs_beagle_stream_close(h)
    return out[:m], samples, sites


def readAD(path, threads=0, rows=None, dtype=np.uint8):
    """The ``--ind_ad_file`` matrix [M, 2N] (plain or gzipped text of integers; ``.npy`` is loaded as is).

    dtype uint8 (default): saturating counts, 255 = "255 reads or more" - the form the GPU upload takes directly
    (sites that deep are never kept by the z-score, zscore.py:36-39); dtype int32: the reference's np.loadtxt matrix."""
    if str(path).endswith(".npy"):
        AD = np.load(path)
        return AD if rows is None else np.ascontiguousarray(AD[rows[0]:rows[1]])
    L = _lib.lib()
    h = ctypes.c_void_p(0)
    t0 = time.perf_counter()
    if L.wgs_ad_stream_open(str(path).encode(), int(threads), ctypes.byref(h)) != 0:
        raise IOError(L.wgs_beagle_last_error().decode())
    try:
        n = L.wgs_ad_stream_inds(h)
        nxt = L.wgs_ad_stream_next_u8 if np.dtype(dtype) == np.uint8 else L.wgs_ad_stream_next_i32
        if rows is not None:
            L.wgs_ad_stream_keep(h, int(rows[0]), int(rows[1]))
            cap = max(int(rows[1]) - int(rows[0]), 0)
        else:
            est = L.wgs_ad_stream_estimate_rows(h)
            cap = int(est * 1.03) + 1024 if est > 0 else 1 << 16
        out = _alloc((max(cap, 1), 2 * n), dtype)
        block = max(1, (64 << 20) // max(2 * n * np.dtype(dtype).itemsize, 1))
        m = 0
        while True:
            if m == cap:
                if rows is not None:
                    break
                cap = int(max(cap * 1.5, L.wgs_ad_stream_estimate_rows(h) * 1.03)) + 1024
                bigger = _alloc((cap, 2 * n), dtype)
                bigger[:m] = out[:m]
                out = bigger
            got = nxt(h, ctypes.c_void_p(out[m:].ctypes.data), min(block, cap - m))
            if got < 0:
                raise IOError(L.wgs_beagle_last_error().decode())
            if got == 0:
                break
            m += got
        _stats(L, h, L.wgs_ad_stream_stats, time.perf_counter() - t0, m, "ad")
    finally:
        L.wgs_ad_stream_close(h)
    return out[:m]
