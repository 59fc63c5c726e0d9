"""Drop-in for the reference's ``reader_cy.readBeagle`` (reader_cy.pyx:16-77), backed by the
multi-threaded C++ parser in ``csrc/wgs_reader.cpp`` (no ``gunzip`` subprocess, no shell
interpolation of the path)."""
import ctypes

import numpy as np

from . import _lib


def readBeagle(beagle, threads=0):
    """Returns ``(L float32 [M, 2N], sample_names list[str], site_names list[str])``."""
    L = _lib.lib()
    h = ctypes.c_void_p(0)
    if L.wgs_beagle_open(str(beagle).encode(), int(threads), ctypes.byref(h)) != 0:
        raise IOError(L.wgs_beagle_last_error().decode())
    try:
        m, n = L.wgs_beagle_sites(h), L.wgs_beagle_inds(h)
        # pinned host memory when a CUDA device is there (full-rate, asynchronous uploads); the parser itself
        # is host code and also runs without one
        try:
            out = _lib.pinned_empty((m, 2 * n), np.float32)
        except _lib.WgsError:
            out = np.empty((m, 2 * n), np.float32)
        L.wgs_beagle_copy(h, ctypes.c_void_p(out.ctypes.data))
        samples = [L.wgs_beagle_sample(h, i).decode() for i in range(n)]
        sites = [L.wgs_beagle_site(h, s).decode() for s in range(m)]
    finally:
        L.wgs_beagle_close(h)
    return out, samples, sites
