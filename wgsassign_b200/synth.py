"""Seeded, depth-consistent synthetic Beagle-style inputs.

The reference ships no allele-depth file and no large data set, so tests and benchmarks use
this generator: per-population allele frequencies, Hardy-Weinberg genotypes, Poisson read
depth, binomial alt reads with sequencing error e, and ANGSD-style genotype likelihoods
rounded to the 6 decimals a Beagle text file carries.  The device-side generator in
``csrc/wgs_kernels.cuh`` (``synth_kernel``) follows the same model with a counter-based hash so that any slice
can be produced in place on the GPU; this NumPy version is for host-sized inputs.
"""
import numpy as np


def make_ids(n_ind, n_pop, interleave=False):
    """IDs array [n,2] of str (sample, population) in Beagle column order.  Individuals
    are split evenly across populations, grouped by population unless `interleave`."""
    if interleave:
        pop_of = np.arange(n_ind) % n_pop
    else:
        pop_of = (np.arange(n_ind) * n_pop) // n_ind
    ids = np.empty((n_ind, 2), dtype="U16")
    for i in range(n_ind):
        ids[i, 0] = "ind%d" % i
        ids[i, 1] = "pop%02d" % pop_of[i]
    return ids


def synth(m, n_ind, n_pop, seed=0, depth=2.0, e=0.01, interleave=False, with_ad=True):
    """Returns dict(L float32 [m,2n], AD int32 [m,2n] (ref, alt), IDs, P float32 [m,K])."""
    rng = np.random.default_rng(seed)
    ids = make_ids(n_ind, n_pop, interleave)
    pops = np.unique(ids[:, 1])
    pop_of = np.searchsorted(pops, ids[:, 1])
    P = np.clip(rng.beta(0.8, 0.8, size=(m, n_pop)), 0.02, 0.98)
    p_ind = P[:, pop_of]                                   # [m, n]
    g = rng.binomial(2, p_ind)                             # alt-allele dosage
    D = rng.poisson(depth, size=(m, n_ind))
    p_alt = np.choose(g, [e, 0.5, 1.0 - e])
    alt = rng.binomial(D, p_alt)
    ref = D - alt
    l0 = (1.0 - e) ** ref * e ** alt
    l1 = 0.5 ** D
    l2 = (1.0 - e) ** alt * e ** ref
    tot = l0 + l1 + l2
    L = np.empty((m, 2 * n_ind), np.float32)
    L[:, 0::2] = np.round(l0 / tot, 6)
    L[:, 1::2] = np.round(l1 / tot, 6)
    out = dict(L=L, IDs=ids, P=P.astype(np.float32), pops=pops)
    if with_ad:
        AD = np.empty((m, 2 * n_ind), np.int32)
        AD[:, 0::2] = ref
        AD[:, 1::2] = alt
        out["AD"] = AD
    return out


def write_beagle(path, L, sample_names, site_names):
    """Write a gzipped Beagle GL text file (format of reader_cy.pyx:31-68): header
    `marker allele1 allele2` + each sample name three times; 6-decimal GLs."""
    import gzip

    m, n = L.shape[0], L.shape[1] // 2
    with gzip.open(path, "wt") as fh:
        fh.write("marker\tallele1\tallele2")
        for s in sample_names:
            fh.write("\t%s\t%s\t%s" % (s, s, s))
        fh.write("\n")
        for r in range(m):
            row = L[r].astype(np.float64)
            g2 = np.round(1.0 - row[0::2] - row[1::2], 6)
            cells = np.empty(3 * n)
            cells[0::3] = row[0::2]
            cells[1::3] = row[1::2]
            cells[2::3] = np.abs(g2)
            fh.write("%s\t0\t1\t" % site_names[r])
            fh.write("\t".join("%.6f" % v for v in cells))
            fh.write("\n")


def bgzf_compress(data, level=1, block=65280):
    """BGZF (bgzip) bytes of `data`: independent gzip members of <= 64 KB of text with their compressed size in a
    "BC" extra subfield, then the end-of-file member - the container ANGSD writes its .beagle.gz in.  For the ingestion
    probe of bench.py and the reader tests."""
    import struct, zlib
    out = bytearray()
    for o in range(0, len(data), block):
        chunk = data[o:o + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(payload) + 8 - 1)
        out += payload + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    return bytes(out) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
