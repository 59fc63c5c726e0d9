"""CPU tests: pin the oracle (oracle/oracle.c + oracle/oracle.py) against the golden
fixtures produced by the unmodified reference CLI, against the reference's own compiled
kernels (oracle/_ref) and, where /root/reference is present, its Python drivers."""
import os

import numpy as np
import pytest

from conftest import parse_tsv


def _backends(oracle):
    b = ["port"]
    if oracle.have_ref():
        b.append("ref")
    return b


def test_port_matches_reference_kernels_bitwise(oracle_mod):
    o = oracle_mod
    if not o.have_ref():
        pytest.skip("oracle/_ref not built")
    from wgsassign_b200 import synth
    d = synth.synth(700, 18, 3, seed=5)
    L, AD, IDs = d["L"], d["AD"], d["IDs"]
    port, ref = o.kernels("port"), o.kernels("ref")
    Lp = np.ascontiguousarray(L[:, o.pop_cols(IDs, "pop01")])
    f1 = np.full(700, 0.25, np.float32); f2 = f1.copy()
    for _ in range(3):
        port.emMAF_update(Lp, f1, 1); ref.emMAF_update(Lp, f2, 2)
    assert np.array_equal(f1, f2)
    assert port.rmse1d(f1, np.full(700, 0.25, np.float32)) == ref.rmse1d(f2, np.full(700, 0.25, np.float32))
    A = np.ascontiguousarray(np.stack([f1, 1 - f1, np.clip(f1 * 1.3, 0, 1)], 1).astype(np.float32))
    v1 = np.zeros(700, np.float32); v2 = v1.copy()
    port.loglike(L, A, v1, 1, 7, 2); ref.loglike(L, A, v2, 1, 7, 2)
    assert np.array_equal(v1, v2)
    a1 = np.zeros(700, np.float32); a2 = a1.copy()
    port.fisher_obs(Lp, A, 1, 1, Lp.shape[1] // 2, a1); ref.fisher_obs(Lp, A, 1, 1, Lp.shape[1] // 2, a2)
    assert np.array_equal(a1, a2, equal_nan=True)
    b1 = np.zeros(700, np.float32); b2 = b1.copy()
    port.ne_obs(a1, A, 1, 1, 6, b1); ref.ne_obs(a2, A, 1, 1, 6, b2)
    assert np.array_equal(b1, b2, equal_nan=True)
    c1 = np.zeros(700, np.float32); c2 = c1.copy()
    port.fisher_obs_ind(L, A, 1, 4, 0, c1); ref.fisher_obs_ind(L, A, 1, 4, 0, c2)
    assert np.array_equal(c1, c2, equal_nan=True)
    # z-score kernels on the restated prep
    summ, arr = o.AD_summary(L, AD, 3, 0, False)
    keep, kept = o.get_L_keep(L, AD, summ, arr, 3)
    fac, like, idx = o.get_factorials(arr, summ, 0.01)
    afv = np.ascontiguousarray(A[keep, 0])
    w1 = np.zeros(kept, np.float32); w2 = w1.copy(); o1 = w1.copy(); o2 = w1.copy()
    port.expected_W_l(L, keep, afv, AD, arr, fac, like, idx, 1, 3, o1, w1)
    ref.expected_W_l(L, keep, afv, AD, arr, fac, like, idx, 1, 3, o2, w2)
    assert np.array_equal(o1, o2) and np.array_equal(w1, w2)
    s1 = np.zeros(kept, np.float32); s2 = s1.copy()
    port.variance_W_l(L, keep, afv, AD, arr, fac, like, idx, 1, 3, s1, w1)
    ref.variance_W_l(L, keep, afv, AD, arr, fac, like, idx, 1, 3, s2, w2)
    assert np.array_equal(s1, s2)


def test_prep_matches_reference_python(oracle_mod):
    """AD_summary / get_L_keep / get_factorials restatements == the reference's zscore.py."""
    o = oracle_mod
    if not (o.have_ref() and os.path.isdir("/root/reference/WGSassign")):
        pytest.skip("reference checkout not present")
    import math
    np.math = math
    o.kernels("ref")
    from WGSassign import zscore as rz
    from wgsassign_b200 import synth
    d = synth.synth(1500, 6, 2, seed=9, depth=3.0)
    L, AD = d["L"], d["AD"]
    for i in (0, 5):
        for thr, single in ((0, False), (10, False), (0, True)):
            s1, a1 = o.AD_summary(L, AD, i, thr, single)
            s2, a2 = rz.AD_summary(L, AD, i, thr, single)
            assert list(s1.keys()) == list(s2.keys())
            for k in s1:
                assert s1[k][0] == s2[k][0] and np.array_equal(s1[k][1], s2[k][1])
            assert np.array_equal(a1, a2)
            k1, n1 = o.get_L_keep(L, AD, s1, a1, i)
            k2, n2 = rz.get_L_keep(L, AD, s2, a2, i)
            assert n1 == n2 and np.array_equal(k1, k2)
            for x, y in zip(o.get_factorials(a1, s1, 0.01), rz.get_factorials(a2, s2, 0.01)):
                assert np.array_equal(x, y)


@pytest.mark.parametrize("backend", ["port", "ref"])
def test_config1_reference_af_and_loo(oracle_mod, bundled, backend):
    o = oracle_mod
    if backend not in _backends(o):
        pytest.skip("backend not built")
    kern = o.kernels(backend)
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    af, pops, its = o.reference_af(L, IDs, 200, 1e-4, 1, kern)
    assert list(pops) == list(bundled["c1_pop_names"])
    assert np.array_equal(af, bundled["c1_pop_af"])
    assert list(its) == list(bundled["c1_em_iters_ref"]) == [17, 14, 16, 14, 13]
    f_obs, ne = o.fisher_obs(L, af, IDs, 1, kern)
    assert np.array_equal(f_obs, bundled["c1_fisher_obs"], equal_nan=True)
    assert np.array_equal(ne, bundled["c1_ne_obs"], equal_nan=True)
    ne_ind = o.fisher_obs_ind(L, af, IDs, 1, kern)
    gold = np.array([float(x) for x in str(bundled["c1_ne_ind_txt"]).split()])
    assert np.array_equal(np.array(["%.7f" % v for v in ne_ind]), np.array(["%.7f" % v for v in gold]))
    ll, parts, its = o.loo(L, af.copy(), IDs, 1, 200, 1e-4, kern=kern)
    assert list(its) == list(bundled["c1_em_iters_loo"])
    hdr, rows = parse_tsv(str(bundled["c1_loo_tsv"]))
    assert hdr == ["sample", "source_pop"] + list(pops)
    got = np.array([["%.6f" % v for v in r] for r in ll])
    want = np.array([r[2:] for r in rows])
    assert np.array_equal(got, want)
    assert rows[0][:2] == ["Ind0", "Northwest"] and rows[0][2] == "-503.560669"


def test_config1_downsampled_partitions(oracle_mod, bundled):
    o = oracle_mod
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    common = np.isin(bundled["sites_breeding"], bundled["sites_breeding_ds"])
    Lf = np.ascontiguousarray(L[common])
    L_ds = bundled["L_breeding_ds"]
    af, pops, _ = o.reference_af(Lf, IDs, 200, 1e-4, 1)
    assert np.array_equal(af, bundled["c1ds_pop_af"])
    ll, parts, its = o.loo(Lf, af.copy(), IDs, 1, 200, 1e-4, downsampled_L=L_ds, num_partitions=3)
    assert list(its) == list(bundled["c1ds_em_iters_loo"])
    _, rows = parse_tsv(str(bundled["c1ds_loo_tsv"]))
    assert np.array_equal(np.array([["%.6f" % v for v in r] for r in ll]), np.array([r[2:] for r in rows]))
    _, prow = parse_tsv(str(bundled["c1ds_loo_parts_tsv"]))
    assert np.array_equal(np.array([["%.6f" % v for v in r] for r in parts]), np.array([r[3:] for r in prow]))


def test_config2_pop_like_and_mixture(oracle_mod, bundled):
    o = oracle_mod
    ll = o.assignLL(bundled["L_nonbreeding"], bundled["c1_pop_af"], 1)
    want = str(bundled["c2_pop_like_txt"]).split()
    assert ["%.7f" % v for v in ll.reshape(-1)] == want
    assert "".join(str(int(x)) for x in np.argmax(ll, 1)) == "1120222122011010110000333013331333"
    mix = o.em_mix(bundled["c2_pop_like"], bundled["IDs_nonbreeding"], 200)
    got = "\n".join(" ".join(str(x) for x in row) for row in mix) + "\n"
    assert got == str(bundled["c2_em_mix_txt"])


@pytest.mark.parametrize("backend", ["port", "ref"])
def test_zscore_golden(oracle_mod, zgold, backend):
    o = oracle_mod
    if backend not in _backends(o):
        pytest.skip("backend not built")
    kern = o.kernels(backend)
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"], zgold["af"]
    pops = np.unique(IDs[:, 1])
    res = o.zscore_assignment(L, AD, af, IDs, pops, kern=kern)
    assert ["%.7f" % r["z"] for r in res] == str(zgold["z_assign_txt"]).split()
    assert [r["loci_kept"] for r in res] == list(zgold["z_assign_loci"])
    comp = zgold["z_assign_components"]
    for i, r in enumerate(res):
        assert np.float32(comp[i, 0]) == r["w_obs"] and np.float32(comp[i, 1]) == r["z_mu"]
        assert np.float32(comp[i, 2]) == r["z_var"]
        assert np.array_equal(r["AD_array"], zgold["AD_array_%d" % i])
        assert np.array_equal(r["L_keep"], zgold["L_keep_%d" % i])
    res = o.zscore_assignment(L, AD, af, IDs, pops, n_threshold=25, ind_start=2, ind_end=9, kern=kern)
    assert ["%.7f" % r["z"] for r in res] == str(zgold["z_assign_thr25_txt"]).split()
    assert [r["loci_kept"] for r in res] == list(zgold["z_assign_thr25_loci"])
    res = o.zscore_assignment(L, AD, af, IDs, pops, single_read=True, kern=kern)
    assert ["%.7f" % r["z"] for r in res] == str(zgold["z_assign_single_txt"]).split()
    res = o.zscore_reference(L, AD, IDs, 200, 1e-4, kern=kern)
    assert ["%.7f" % r["z"] for r in res] == str(zgold["z_ref_txt"]).split()
    assert [r["loci_kept"] for r in res] == list(zgold["z_ref_loci"])
    assert [r["em_iter"] for r in res] == list(zgold["z_ref_em_iters"])
