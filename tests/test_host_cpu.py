"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), the C++ Beagle reader, the file writers, the CLI
surface, and the site-sharding plumbing under gloo."""
import gzip
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from wgsassign_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    return _lib


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "wgsassign_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wgs_[a-z_A-Z0-9]+)\s*\(", hdr)) - {"wgs_allreduce_fn"}
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.wgs_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, out), name


def test_no_cpu_fallback(lib):
    if lib.lib().wgs_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(lib.WgsError, match="no CPU fallback"):
        lib.Context(0)
    from wgsassign_b200 import glassy
    with pytest.raises(lib.WgsError):
        glassy.assignLL(np.zeros((4, 4), np.float32), np.zeros((4, 1), np.float32), 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "wgsassign_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f in ("synth.py",) or "import oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_reader_roundtrip_and_formats(lib, tmp_path):
    from wgsassign_b200 import reader, synth
    d = synth.synth(257, 7, 2, seed=4, with_ad=False)
    L = d["L"]
    sites = ["chr%d_%d" % (1 + s % 3, 10 * s + 5) for s in range(257)]
    names = ["s%02d" % i for i in range(7)]
    p = str(tmp_path / "a.beagle.gz")
    synth.write_beagle(p, L, names, sites)
    for t in (1, 3):
        L2, n2, s2 = reader.readBeagle(p, threads=t)
        assert L2.dtype == np.float32 and np.array_equal(L2, L) and n2 == names and s2 == sites
    # mixed delimiters, scientific notation, many digits, CRLF, blank trailing line
    p2 = str(tmp_path / "b.beagle.gz")
    with gzip.open(p2, "wt") as fh:
        fh.write("marker allele1\tallele2 A A A  B B B\r\n")
        fh.write("c_1 0 1 3.3e-1 0.33333333333333333333 0.3   1 0 0\r\n")
        fh.write("c_2\t2\t3\t0.000001\t.5\t0.499999\t1e-7 0.25 0.75\n\n")
    L3, n3, s3 = reader.readBeagle(p2)
    want = np.array([[np.float32(float("3.3e-1")), np.float32(float("0.33333333333333333333")), 1, 0],
                     [np.float32(0.000001), np.float32(.5), np.float32(1e-7), np.float32(0.25)]], np.float32)
    assert n3 == ["A", "B"] and s3 == ["c_1", "c_2"] and np.array_equal(L3, want)
    # a short row is an error, not undefined behaviour
    p3 = str(tmp_path / "c.beagle.gz")
    with gzip.open(p3, "wt") as fh:
        fh.write("marker allele1 allele2 A A A\nc_1 0 1 0.3 0.3\n")
    with pytest.raises(IOError):
        reader.readBeagle(p3)
    with pytest.raises(IOError):
        reader.readBeagle(str(tmp_path / "missing.gz"))


def test_reader_row_ranges_blocks_and_whole_file_form(lib, tmp_path):
    """A site-sharded rank converts only its own rows (names and the row count still cover the whole file); blocks
    are reported as they complete; the one-call C form gives the same matrix; files spanning several inflate
    blocks (> 32 MB of text) keep their row order."""
    import ctypes
    from wgsassign_b200 import reader, synth
    m, n = 3001, 9
    d = synth.synth(m, n, 3, seed=5, with_ad=False)
    L = d["L"]
    sites = ["c%d_%d" % (s % 5, s) for s in range(m)]
    names = ["i%d" % i for i in range(n)]
    p = str(tmp_path / "r.beagle.gz")
    synth.write_beagle(p, L, names, sites)
    mm, nn, ss = reader.count_rows(p, threads=2)
    assert mm == m and nn == names and ss == sites
    for lo, hi in ((0, 1000), (1000, 2000), (2000, m), (m, m), (17, 18)):
        seen = []
        Lr, n2, s2 = reader.readBeagle(p, threads=3, rows=(lo, hi), on_block=lambda a, r0, r1: seen.append((r0, r1)))
        assert Lr.shape == (hi - lo, 2 * n) and np.array_equal(Lr, L[lo:hi]) and n2 == names and s2 == sites
        assert (not seen and hi == lo) or (seen[0][0] == 0 and seen[-1][1] == hi - lo)
    Lc = lib.lib()
    h = ctypes.c_void_p(0)
    assert Lc.wgs_beagle_open(p.encode(), 2, ctypes.byref(h)) == 0
    out = np.empty((Lc.wgs_beagle_sites(h), 2 * Lc.wgs_beagle_inds(h)), np.float32)
    Lc.wgs_beagle_copy(h, ctypes.c_void_p(out.ctypes.data))
    assert Lc.wgs_beagle_site(h, m - 1).decode() == sites[-1]
    Lc.wgs_beagle_close(h)
    assert np.array_equal(out, L)
    # ~40 MB of text: two inflate blocks, rows split across the block boundary
    m2, n2_ = 2600, 600
    big = np.round(np.random.default_rng(3).random((m2, 2 * n2_)) * 0.5, 6).astype(np.float32)
    p2 = str(tmp_path / "big.beagle.gz")
    synth.write_beagle(p2, big, ["s%d" % i for i in range(n2_)], ["x_%d" % s for s in range(m2)])
    Lb, _, sb = reader.readBeagle(p2, threads=4)
    assert np.array_equal(Lb, big) and sb[-1] == "x_%d" % (m2 - 1)
    assert reader.last_stats["beagle"]["uncompressed_bytes"] > (32 << 20)


def test_allele_depth_reader(lib, tmp_path):
    """--ind_ad_file: plain or gzipped integer text -> the reference's int32 matrix (np.loadtxt) or saturating uint8."""
    from wgsassign_b200 import reader
    rng = np.random.default_rng(8)
    AD = rng.poisson(2.0, size=(1234, 2 * 11)).astype(np.int32)
    AD[5, 3], AD[9, 0] = 255, 1000
    p = str(tmp_path / "ad.txt")
    np.savetxt(p, AD, fmt="%d")
    pz = str(tmp_path / "ad.txt.gz")
    with gzip.open(pz, "wt") as fh:
        fh.write(open(p).read())
    ref = np.loadtxt(p, dtype=np.int32)
    for path in (p, pz):
        a32 = reader.readAD(path, threads=3, dtype=np.int32)
        assert a32.dtype == np.int32 and np.array_equal(a32, ref)
        a8 = reader.readAD(path, threads=2)
        assert a8.dtype == np.uint8 and np.array_equal(a8, np.minimum(ref, 255).astype(np.uint8))
        part = reader.readAD(path, rows=(100, 300))
        assert np.array_equal(part, np.minimum(ref[100:300], 255).astype(np.uint8))
    np.save(str(tmp_path / "ad.npy"), AD)
    assert np.array_equal(reader.readAD(str(tmp_path / "ad.npy")), AD)
    bad = str(tmp_path / "bad.txt")
    open(bad, "w").write("1 2 3 4\n1 2 3\n")
    with pytest.raises(IOError):
        reader.readAD(bad)
    open(bad, "w").write("1 2 3\n")
    with pytest.raises(IOError, match="two counts"):
        reader.readAD(bad)


def test_reader_matches_reference_reader_on_bundled_files(lib, bundled):
    ref = "/root/reference/data/amre.breeding.ind85.ds_2x.sites-filter.top_50_each.beagle.gz"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present")
    from wgsassign_b200 import reader
    L, names, sites = reader.readBeagle(ref)
    assert np.array_equal(L, bundled["L_breeding"])
    assert names == list(bundled["samples_breeding"]) and sites == list(bundled["sites_breeding"])


def test_writers_and_mixture(bundled, tmp_path):
    from wgsassign_b200 import mixture, utils
    mix = mixture.em_mix(bundled["c2_pop_like"], bundled["IDs_nonbreeding"], 200)
    assert "\n".join(" ".join(str(x) for x in row) for row in mix) + "\n" == str(bundled["c2_em_mix_txt"])
    # LOO TSV format: feed the golden numbers back through our writer
    hdr_rows = str(bundled["c1_loo_tsv"]).strip().split("\n")
    vals = np.array([[float(x) for x in r.split("\t")[2:]] for r in hdr_rows[1:]], np.float32)
    out = str(tmp_path / "x.tsv")
    utils.write_ass_mats(out, vals, list(bundled["samples_breeding"]), bundled["c1_pop_names"], print_part_column=False,
                         sample_locations=bundled["IDs_breeding"][:, 1], doing_LOO=True)
    assert open(out).read() == str(bundled["c1_loo_tsv"])
    with pytest.raises(ValueError):
        utils.write_ass_mats(out, vals[:3], list(bundled["samples_breeding"]), bundled["c1_pop_names"])
    v = np.arange(10, dtype=np.float32)
    assert np.array_equal(utils.partition_loglikes(v, 3), np.array([0 + 3 + 6 + 9, 1 + 4 + 7, 2 + 5 + 8], np.float32))


def test_em_mix_logsumexp_option(bundled):
    """--em_mix_logsumexp: the same proportions as the reference's arithmetic where that is finite, finite ones where it is
    not (log-likelihoods of a genome-scale run underflow exp())."""
    from wgsassign_b200 import mixture
    rng = np.random.default_rng(2)
    ll = -(rng.random((30, 4)) * 40 + 400)                      # bundled-data scale: exp() is finite
    idx = np.stack([np.arange(30).astype("U4"), np.array(["a", "b", "c"])[np.arange(30) % 3]], axis=1)
    a = mixture.em_mix(ll, idx, 200)
    b = mixture.em_mix(ll, idx, 200, logsumexp=True)
    assert np.array_equal(a[:, 0], b[:, 0]) and np.allclose(a[:, 1:].astype(float), b[:, 1:].astype(float), rtol=1e-5, atol=1e-7)
    big = ll * 1e4                                              # genome scale: the reference's form is NaN, the hardened one is not
    with np.errstate(all="ignore"):
        assert np.all(np.isnan(mixture.em_mix(big, idx, 50)[:, 1:].astype(float)))
    c = mixture.em_mix(big, idx, 50, logsumexp=True)[:, 1:].astype(float)
    assert np.all(np.isfinite(c)) and np.allclose(c.sum(1), 1.0, atol=1e-5)


def test_cli_surface():
    from wgsassign_b200 import WGSassign as cli
    flags = {a.dest: a.default for a in cli.parser._actions if a.dest != "help"}
    expect = {"beagle": None, "threads": 1, "out": "wgsassign", "maf_iter": 200, "maf_tole": 1e-4, "pop_af_IDs": None,
              "get_reference_af": False, "pop_names": None, "ne_obs": False, "loo": False, "loo_downsampled_beagle": None,
              "pop_af_file": None, "get_pop_like": False, "partition_sites": 1, "get_assignment_z_score": False,
              "get_reference_z_score": False, "ind_ad_file": None, "allele_count_threshold": None,
              "single_read_threshold": False, "ind_start": None, "ind_end": None, "pop_like": None, "pop_like_IDs": None,
              "get_em_mix": False, "get_mcmc_mix": False, "mixture_iter": 200}
    extra = {"em_mix_logsumexp": False, "shard_by_bytes": False}          # additions of this implementation: optional, default = the reference's behaviour
    assert flags == {**expect, **extra}
    ref_cli = "/root/reference/WGSassign/WGSassign.py"
    if os.path.exists(ref_cli):
        ref_flags = set(re.findall(r'add_argument\(\s*(?:"-\w",\s*)?["\']--(\w+)["\']', open(ref_cli).read()))
        ref_flags -= {"plink", "iter", "tole"}      # commented out in the reference
        assert ref_flags == set(expect), ref_flags ^ set(expect)
    with pytest.raises(ValueError, match="requires that --loo"):
        cli.main(["--loo_downsampled_beagle", "x.gz"])


def test_cli_mixture_mode_end_to_end(bundled, tmp_path):
    """--get_em_mix needs no GPU: byte-compare with the reference CLI's output."""
    from wgsassign_b200 import WGSassign as cli
    pl = str(tmp_path / "pl.txt")
    open(pl, "w").write(str(bundled["c2_pop_like_txt"]))
    ids = str(tmp_path / "ids.txt")
    np.savetxt(ids, bundled["IDs_nonbreeding"], fmt="%s", delimiter="\t")
    out = str(tmp_path / "o")
    cli.main(["--pop_like", pl, "--pop_like_IDs", ids, "--get_em_mix", "--out", out])
    assert open(out + ".em_mix.txt").read() == str(bundled["c2_em_mix_txt"])
    assert "-get_em_mix" in open(out + ".args").read()


def test_shard_ranges():
    from wgsassign_b200 import dist
    for M in (0, 1, 7, 449, 10 ** 6 + 3):
        for w in (1, 2, 3, 8):
            r = [dist.shard_range(M, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == M and all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_sharded_stop_rule_matches_unsharded(oracle_mod, bundled):
    """Design check for SURVEY 8e.4: summing per-shard squared changes reproduces the global
    stop iteration; stopping each shard on its own RMSE does not."""
    o = oracle_mod
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    Lp = np.ascontiguousarray(L[:, o.pop_cols(IDs, "Northwest")])
    kern = o.kernels("port")
    f_ref, it_ref = o.emMAF(Lp, 200, 1e-4, 1, kern)
    shards = [np.ascontiguousarray(Lp[a:b]) for a, b in ((0, 200), (200, 449))]
    fs = [np.full(s.shape[0], 0.25, np.float32) for s in shards]
    it_glob = 0
    for it in range(1, 201):
        ssq = 0.0
        for s, f in zip(shards, fs):
            prev = f.copy()
            kern.emMAF_update(s, f, 1)
            ssq += float(np.sum((f.astype(np.float64) - prev) ** 2))
        if np.sqrt(np.float32(np.float32(ssq) / np.float32(449))) < 1e-4:
            it_glob = it
            break
    assert it_glob == it_ref and np.array_equal(np.concatenate(fs), f_ref)
    own = [o.emMAF(s, 200, 1e-4, 1, kern)[1] for s in shards]
    assert own != [it_ref, it_ref]


GLOO_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as td
from wgsassign_b200 import dist
td.init_process_group("gloo")
r, w = td.get_rank(), td.get_world_size()
M = 11
lo, hi = dist.shard_range(M, r, w)
dist.enable(M, lo)
assert dist.enabled() and dist.total_sites(hi - lo) == M
a = np.arange(6, dtype=np.float64) * (r + 1)
dist.allreduce_sum(a)
assert np.array_equal(a, np.arange(6) * sum(range(1, w + 1)))
i = np.array([r + 1, 10 * (r + 1)], dtype=np.int64)
dist.allreduce_sum(i)
assert list(i) == [sum(range(1, w + 1)), 10 * sum(range(1, w + 1))]
rows = np.arange(lo, hi, dtype=np.float32).reshape(-1, 1) * np.ones((1, 3), np.float32)
full = dist.gather_rows(rows)
assert full.shape == (M, 3) and np.array_equal(full[:, 0], np.arange(M))
td.destroy_process_group()
print("ok", r)
"""


def test_dist_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2


def test_numa_binding_helper_is_failsafe(tmp_path, monkeypatch):
    """dist.bind_near_gpu: parses sysfs cpulists, never raises, and leaves the affinity alone when there is no GPU."""
    import os
    from wgsassign_b200 import dist
    assert dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert dist._parse_cpulist("5") == {5}
    before = os.sched_getaffinity(0)
    msg = dist.bind_near_gpu(0, sysfs=str(tmp_path))            # no device (CPU box) or no such sysfs tree: a message, no change
    assert msg.startswith("numa:")
    assert os.sched_getaffinity(0) == before
    monkeypatch.setenv("WGS_NO_NUMA_BIND", "1")
    assert "off" in dist.bind_near_gpu(0)


def test_bgzf_input_reads_like_plain_gzip(tmp_path):
    """A BGZF Beagle / allele-depth file (members inflated in parallel) gives exactly what the same text gives as one
    gzip stream: values, names, row ranges; a truncated or corrupted member is an error, not silent data."""
    import gzip
    from wgsassign_b200 import reader, synth
    n, m = 7, 9000                                               # ~ 2 MB of text: dozens of 64 KB members, lines straddle them
    rng = np.random.default_rng(5)
    vals = np.round(rng.random((m, 3 * n)), 6)
    header = "marker\tallele1\tallele2" + "".join("\tind%d\tind%d\tind%d" % (i, i, i) for i in range(n)) + "\n"
    text = (header + "".join("chr1_%d\t0\t1\t%s\n" % (s, "\t".join("%.6f" % v for v in vals[s])) for s in range(m))).encode()
    pg, pz = tmp_path / "a.beagle.gz", tmp_path / "a.bgzf.beagle.gz"
    pg.write_bytes(gzip.compress(text, 1))
    z = synth.bgzf_compress(text)
    assert gzip.decompress(z) == text
    pz.write_bytes(z)
    A = reader.readBeagle(str(pg), 4)
    assert reader.last_stats["beagle"]["bgzf"] is False
    B = reader.readBeagle(str(pz), 4)
    assert reader.last_stats["beagle"]["bgzf"] is True
    assert np.array_equal(A[0], B[0]) and list(A[1]) == list(B[1]) and list(A[2]) == list(B[2])
    assert A[0].shape == (m, 2 * n)
    lo, hi = 1234, 7777
    C = reader.readBeagle(str(pz), 3, rows=(lo, hi))
    assert np.array_equal(C[0], A[0][lo:hi])
    assert reader.count_rows(str(pz))[0] == m
    # allele depths
    ad = rng.integers(0, 40, size=(3000, 2 * n))
    atext = "".join(" ".join(map(str, r)) + "\n" for r in ad).encode()
    qa, qz = tmp_path / "ad.txt.gz", tmp_path / "ad.bgzf.txt.gz"
    qa.write_bytes(gzip.compress(atext, 1))
    qz.write_bytes(synth.bgzf_compress(atext))
    assert np.array_equal(reader.readAD(str(qa), 4), reader.readAD(str(qz), 4))
    # damage: flip a payload byte of a middle member / cut the file inside a member
    bad = bytearray(z)
    bad[len(bad) // 2] ^= 0x55
    (tmp_path / "bad.beagle.gz").write_bytes(bytes(bad))
    with pytest.raises(Exception):
        reader.readBeagle(str(tmp_path / "bad.beagle.gz"), 4)
    (tmp_path / "cut.beagle.gz").write_bytes(z[: len(z) // 2])
    with pytest.raises(Exception):
        reader.readBeagle(str(tmp_path / "cut.beagle.gz"), 4)


def test_reader_tokens_equal_float_of_atof(tmp_path):
    """Every token form around the fixed-point fast path converts to (float)atof(token): random D.DDDDDD values
    (the path itself), eight-character tokens that are NOT of that form, and longer / shorter / signed / exponent forms."""
    import gzip
    from wgsassign_b200 import reader
    rng = np.random.default_rng(11)
    toks = ["%.6f" % v for v in rng.random(4000)] + ["%.6f" % v for v in (0.0, 1.0, 0.333333, 0.666667, 0.999999, 0.000001, 9.999999)]
    toks += ["1.00e-05", "-0.12345", "12.34567", "+1.23456", "1234567.", ".1234567", "1e-00005", "0.12345e", "00000001", "0.5", "1", "0",
             "0.1234567", "0.12345678901234567", "3.0e-1", "1E-3", "-0.000000", "0.0000001", "123456.5", "1.0000000"]
    n = 3
    per_row = 2 * n                                                  # tokens that are kept per row (the third of each triple is dropped)
    while len(toks) % per_row:
        toks.append("0.250000")
    rows = []
    for r in range(len(toks) // per_row):
        t = toks[r * per_row:(r + 1) * per_row]
        cells = []
        for i in range(n):
            cells += [t[2 * i], t[2 * i + 1], "0.111111"]
        rows.append("s_%d\t0\t1\t%s\n" % (r, "\t".join(cells)))
    header = "marker\tallele1\tallele2" + "".join("\ti%d\ti%d\ti%d" % (i, i, i) for i in range(n)) + "\n"
    path = tmp_path / "tok.beagle.gz"
    path.write_bytes(gzip.compress((header + "".join(rows)).encode(), 1))
    L, _, _ = reader.readBeagle(str(path), 2)
    import ctypes, ctypes.util
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.atof.restype = ctypes.c_double
    want = np.array([np.float32(libc.atof(t.encode())) for t in toks], np.float32).reshape(-1, per_row)
    assert L.shape == want.shape
    assert np.array_equal(L.view(np.uint32), want.view(np.uint32))


GLOO_PARTS_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as td
from wgsassign_b200 import dist, reader
td.init_process_group("gloo")
r, w = td.get_rank(), td.get_world_size()
path = %r
L, samples, sites = reader.readBeagle(path, 2, part=(r, w))
M, lo, ranges = dist.row_counts_to_ranges(L.shape[0])
dist.enable(M, lo, ranges=ranges)
assert ranges[r] == (lo, lo + L.shape[0]) and dist.rank_ranges() == ranges
full = dist.gather_rows(L)                       # unequal row blocks, concatenated in rank order
whole, wsamples, wsites = reader.readBeagle(path, 2)
assert M == whole.shape[0] and np.array_equal(full, whole) and list(samples) == list(wsamples)
assert list(sites) == list(wsites)[lo:lo + L.shape[0]]
try:
    dist.enable(M, lo + 1, ranges=ranges)
    raise SystemExit("inconsistent ranges accepted")
except ValueError:
    pass
td.destroy_process_group()
print("ok", r, L.shape[0])
"""


def test_byte_range_parts_world3_gloo(tmp_path):
    """Three processes read the three byte ranges of one BGZF file: their row counts give the shard geometry
    (dist.row_counts_to_ranges / enable(ranges=...)), and the gathered rows are the whole file."""
    from wgsassign_b200 import synth
    n, m = 5, 4000
    rng = np.random.default_rng(9)
    vals = np.round(rng.random((m, 3 * n)), 6)
    header = "marker\tallele1\tallele2" + "".join("\ti%d\ti%d\ti%d" % (i, i, i) for i in range(n)) + "\n"
    text = (header + "".join("c_%d\t0\t1\t%s\n" % (s, "\t".join("%.6f" % v for v in vals[s])) for s in range(m))).encode()
    path = tmp_path / "p.beagle.gz"
    path.write_bytes(synth.bgzf_compress(text, block=20000))
    script = tmp_path / "w.py"
    script.write_text(GLOO_PARTS_WORKER % (ROOT, str(path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 3


def test_byte_range_parts_cover_the_file(tmp_path):
    """reader.readBeagle(part=(p, P)): for every P the parts are disjoint, ordered and cover the file - rows longer than a
    BGZF member, members far smaller than a row, more parts than members, a file without the end-of-file member."""
    import gzip
    from wgsassign_b200 import reader, synth
    rng = np.random.default_rng(3)
    for n, m, block, strip_eof in [(7, 3000, 65280, False), (3, 50, 65280, True), (2000, 12, 65280, False), (5, 1500, 777, True), (1, 5, 65280, False)]:
        vals = np.round(rng.random((m, 3 * n)), 6)
        header = "marker\tallele1\tallele2" + "".join("\ti%d\ti%d\ti%d" % (i, i, i) for i in range(n)) + "\n"
        text = (header + "".join("c_%d\t0\t1\t%s\n" % (s, "\t".join("%.6f" % v for v in vals[s])) for s in range(m))).encode()
        z = synth.bgzf_compress(text, block=block)
        path = tmp_path / ("x_%d_%d.beagle.gz" % (n, block))
        path.write_bytes(z[:-28] if strip_eof else z)
        assert reader.is_bgzf(str(path))
        A = reader.readBeagle(str(path), 3)
        for P in (2, 3, 8, 64):
            parts = [reader.readBeagle(str(path), 2, part=(q, P)) for q in range(P)]
            assert np.array_equal(np.concatenate([x[0] for x in parts], axis=0), A[0]), (n, m, block, P)
            assert sum([list(x[2]) for x in parts], []) == list(A[2])
            assert all(list(x[1]) == list(A[1]) for x in parts)
    plain = tmp_path / "g.beagle.gz"
    plain.write_bytes(gzip.compress(b"marker\ta\tb\ti\ti\ti\nc\t0\t1\t0.1\t0.2\t0.7\n"))
    assert not reader.is_bgzf(str(plain))
    with pytest.raises(IOError):
        reader.readBeagle(str(plain), 2, part=(0, 2))


GLOO_CLI_PARTS_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as td
from wgsassign_b200 import WGSassign, dist, reader, session

class FakeCtx:                                   # stands in for the device context: records what the host layer tells it
    def set_shard(self, M_total, offset, fn): self.shard = (M_total, offset)
    def set_rank(self, rank, world): self.rank = (rank, world)

fake = FakeCtx()
def stream_context(beagle, pop_of_ind=None, K=0, threads=0, rows=None, part=None):
    assert rows is None and part == (td.get_rank(), td.get_world_size())
    L, samples, sites = reader.readBeagle(beagle, threads, part=part)
    return fake, L, samples, sites
session.stream_context = stream_context

path = %r
args = WGSassign.parser.parse_args(["--beagle", path, "--shard_by_bytes", "--get_pop_like", "--threads", "2"])
run = WGSassign._Run(args)                        # initialises torch.distributed (gloo: no CUDA here)
run.parse_inputs()
r, w = td.get_rank(), td.get_world_size()
whole, _, wsites = reader.readBeagle(path, 2)
lo, hi = run._ranges[r]
assert run.M_total == whole.shape[0] and hi - lo == run.L.shape[0]
assert np.array_equal(run.L, whole[lo:hi]) and list(run.site_names) == list(wsites)[lo:hi]
assert fake.shard == (run.M_total, lo) and fake.rank == (r, w)
assert run._range(run.M_total) == (lo, hi)
X = np.arange(run.M_total * 2, dtype=np.float32).reshape(run.M_total, 2)      # a per-site input read whole (e.g. --pop_af_file)
assert np.array_equal(run.shard(X), X[lo:hi])
assert np.array_equal(run.full(run.shard(X)), X)
td.destroy_process_group()
print("ok", r, hi - lo)
"""


def test_cli_shard_by_bytes_host_logic_gloo(tmp_path):
    """--shard_by_bytes under torchrun (gloo, two processes): parse_inputs reads each rank's byte range, derives the
    shard geometry from the row counts, attaches the context with it, and shard() / full() follow the unequal ranges."""
    from wgsassign_b200 import synth
    n, m = 4, 3000
    rng = np.random.default_rng(13)
    vals = np.round(rng.random((m, 3 * n)), 6)
    header = "marker\tallele1\tallele2" + "".join("\ti%d\ti%d\ti%d" % (i, i, i) for i in range(n)) + "\n"
    text = (header + "".join("c_%d\t0\t1\t%s\n" % (s, "\t".join("%.6f" % v for v in vals[s])) for s in range(m))).encode()
    path = tmp_path / "p.beagle.gz"
    path.write_bytes(synth.bgzf_compress(text, block=30000))
    script = tmp_path / "w.py"
    script.write_text(GLOO_CLI_PARTS_WORKER % (ROOT, str(path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29615", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2
