"""GPU parity of the z-score path against the golden fixture produced by the reference CLI
(tests/golden/zscore.npz) and against the oracle on other seeds.

Bit-exact: loci kept, allele-depth class tallies (AD_array rows), EM iteration counts.
1e-6 relative: the three float32 components (W_obs, z_mu, z_var).  z itself is the
difference of two sums of magnitude ~|W| divided by sqrt(var), so its tolerance is the
propagated component tolerance (hard part 4).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zs():
    import wgsassign_b200._lib as _lib
    if _lib.lib().wgs_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    from wgsassign_b200 import session, zscore
    session.reset()
    yield zscore
    session.reset()


def sort_rows(a):
    a = np.asarray(a).reshape(-1, 4)
    return a[np.lexsort((a[:, 1], a[:, 2]))]


def check_rows(got, ref_z, ref_comp, ref_loci, ref_classes=None):
    for j, r in enumerate(got):
        assert r["loci_kept"] == ref_loci[j]
        w, mu, var = ref_comp[j]
        for a, b in ((r["w_obs"], w), (r["z_mu"], mu), (r["z_var"], var)):
            assert abs(float(a) - float(b)) <= 1e-6 * abs(float(b)), (j, a, b)
        tol = 2e-6 * (abs(w) + abs(mu)) / np.sqrt(var) + 1e-5 * abs(ref_z[j])
        assert abs(float(r["z"]) - ref_z[j]) <= tol, (j, r["z"], ref_z[j], tol)
        if ref_classes is not None:
            assert np.array_equal(sort_rows(r["AD_array"]), sort_rows(ref_classes[j]))


def test_assignment_zscore_golden(zs, zgold):
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"], zgold["af"]
    pops = np.unique(IDs[:, 1])
    got = zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops)
    ref_z = [float(x) for x in str(zgold["z_assign_txt"]).split()]
    comp = zgold["z_assign_components"]
    check_rows(got, ref_z, comp[:, :3], list(zgold["z_assign_loci"]), [zgold["AD_array_%d" % i] for i in range(12)])
    assert [r["loci_kept"] for r in got] == [int(x) for x in comp[:, 3]]


def test_assignment_zscore_thresholds(zs, zgold, oracle_mod):
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"], zgold["af"]
    pops = np.unique(IDs[:, 1])
    got = zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops, n_threshold=25, ind_start=2, ind_end=9)
    ref = oracle_mod.zscore_assignment(L, AD, af, IDs, pops, n_threshold=25, ind_start=2, ind_end=9)
    assert ["%.7f" % r["z"] for r in ref] == str(zgold["z_assign_thr25_txt"]).split()
    check_rows(got, [float(r["z"]) for r in ref], [(float(r["w_obs"]), float(r["z_mu"]), float(r["z_var"])) for r in ref],
               list(zgold["z_assign_thr25_loci"]), [r["AD_array"] for r in ref])
    got = zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops, single_read=True)
    ref = oracle_mod.zscore_assignment(L, AD, af, IDs, pops, single_read=True)
    check_rows(got, [float(r["z"]) for r in ref], [(float(r["w_obs"]), float(r["z_mu"]), float(r["z_var"])) for r in ref],
               list(zgold["z_assign_single_loci"]), [r["AD_array"] for r in ref])


def test_reference_zscore_golden(zs, zgold, oracle_mod):
    L, AD, IDs = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"]
    got = zs.zscore_all(L, AD, IDs, zs.MODE_REFERENCE)
    ref = oracle_mod.zscore_reference(L, AD, IDs, 200, 1e-4)
    assert ["%.7f" % r["z"] for r in ref] == str(zgold["z_ref_txt"]).split()
    assert [r["em_iters"] for r in got] == list(zgold["z_ref_em_iters"])
    check_rows(got, [float(r["z"]) for r in ref], [(float(r["w_obs"]), float(r["z_mu"]), float(r["z_var"])) for r in ref],
               list(zgold["z_ref_loci"]), [r["AD_array"] for r in ref])


@pytest.mark.parametrize("m,n,k,depth,seed", [(5000, 40, 4, 2.0, 31), (3000, 21, 3, 5.0, 32)])
def test_zscore_vs_oracle_synthetic(zs, oracle_mod, m, n, k, depth, seed):
    from wgsassign_b200 import synth
    d = synth.synth(m, n, k, seed=seed, depth=depth, interleave=True)
    L, AD, IDs = d["L"], d["AD"], d["IDs"]
    af, pops, _ = oracle_mod.reference_af(L, IDs, 200, 1e-4, 2)
    sub = (3, min(n, 11))
    got = zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops, ind_start=sub[0], ind_end=sub[1])
    ref = oracle_mod.zscore_assignment(L, AD, af, IDs, pops, ind_start=sub[0], ind_end=sub[1], t=2)
    check_rows(got, [float(r["z"]) for r in ref], [(float(r["w_obs"]), float(r["z_mu"]), float(r["z_var"])) for r in ref],
               [r["loci_kept"] for r in ref], [r["AD_array"] for r in ref])
    got = zs.zscore_all(L, AD, IDs, zs.MODE_REFERENCE, ind_start=sub[0], ind_end=sub[1])
    ref = oracle_mod.zscore_reference(L, AD, IDs, 200, 1e-4, ind_start=sub[0], ind_end=sub[1], t=2)
    assert [r["em_iters"] for r in got] == [r["em_iter"] for r in ref]
    check_rows(got, [float(r["z"]) for r in ref], [(float(r["w_obs"]), float(r["z_mu"]), float(r["z_var"])) for r in ref],
               [r["loci_kept"] for r in ref], [r["AD_array"] for r in ref])


def test_zscore_errors(zs, zgold):
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"], zgold["af"]
    pops = np.unique(IDs[:, 1])
    with pytest.raises(AssertionError, match="loci were kept"):
        zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops, n_threshold=10 ** 6)
    neg = AD.copy()
    neg[0, 0] = -1
    with pytest.raises(Exception, match="negative"):
        zs.zscore_all(L, neg, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops)
    with pytest.raises(ValueError, match="af must be"):
        zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=np.ascontiguousarray(af[:-3]), pops=pops)


def test_extreme_depths_do_not_abort(zs, zgold, oracle_mod):
    """A few sites with hundreds of reads (common in real depth files): the reference gives each its own class,
    which is never kept; here they are stored as "deeper than the class table".  Same tallies, kept loci and scores
    as the oracle on the same matrix - for the individual that has them and for everybody else."""
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32).copy(), zgold["IDs"], zgold["af"]
    pops = np.unique(IDs[:, 1])
    AD[3, 0], AD[3, 1] = 300, 2            # individual 0
    AD[7, 2], AD[7, 3] = 90, 1000          # individual 1
    AD[9, 0], AD[9, 1] = 30, 25            # depth 55: above the table's cap of 40, below the uint8 limit
    got = zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops, ind_start=0, ind_end=4)
    ref = oracle_mod.zscore_assignment(L, AD, af, IDs, pops, ind_start=0, ind_end=4)
    check_rows(got, [float(r["z"]) for r in ref], [(float(r["w_obs"]), float(r["z_mu"]), float(r["z_var"])) for r in ref],
               [r["loci_kept"] for r in ref], [r["AD_array"] for r in ref])
    from wgsassign_b200 import session
    assert session._state["ctx"].zscore_deep_sites() == 3


def test_exact_mean_variant_keeps_tallies(zs, zgold):
    """The order-independent fixed-point tally (option z_exact_means, NOT the default): identical integer
    tallies; class means differ from numpy's sequential float32 mean only by that mean's own
    accumulation error (4e-6 relative at 400 sites), which can move a site sitting on the 0.01
    keep threshold, so loci kept agree to a couple of sites and the sums to ~1e-3."""
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"], zgold["af"]
    pops = np.unique(IDs[:, 1])
    from wgsassign_b200 import session
    session.set_option("z_exact_means", 1)
    try:
        got = zs.zscore_all(L, AD, IDs, zs.MODE_ASSIGNMENT, A=af, pops=pops)
    finally:
        session.set_option("z_exact_means", 0)
    comp = zgold["z_assign_components"]
    for j, r in enumerate(got):
        assert abs(r["loci_kept"] - int(comp[j, 3])) <= 3
        assert np.array_equal(sort_rows(r["AD_array"]), sort_rows(zgold["AD_array_%d" % j]))
        assert abs(float(r["w_obs"]) - comp[j, 0]) <= 2e-3 * abs(comp[j, 0])
        assert abs(float(r["z_mu"]) - comp[j, 1]) <= 2e-3 * abs(comp[j, 1])


def test_per_individual_functions_match_reference_loop(zs, zgold, oracle_mod):
    """The reference's own per-individual body (WGSassign.py:425-443) written against our
    drop-in functions: dict keys in first-occurrence order, AD_array / L_keep / AD_index
    bit-exact, class means and per-site W arrays within float32 rounding of the oracle."""
    L, AD, IDs, af = zgold["L"], zgold["AD"].astype(np.int32), zgold["IDs"], zgold["af"]
    pops = np.unique(IDs[:, 1])
    kern = oracle_mod.kernels()
    for i in (0, 5, 11):
        k = int(np.argwhere(pops == IDs[i, 1])[0][0])
        d, arr = zs.AD_summary(L, AD, i, 0, False)
        d_o, arr_o = oracle_mod.AD_summary(L, AD, i, 0, False)
        assert [tuple(int(x) for x in key) for key in d] == list(d_o)
        assert arr.dtype == np.int32 and np.array_equal(arr, arr_o)
        for key, key_o in zip(d, d_o):
            assert d[key][0] == d_o[key_o][0]
            assert d[key][1].dtype == np.float32 and np.allclose(d[key][1], d_o[key_o][1], rtol=1e-6, atol=1e-9)
        keep, kept = zs.get_L_keep(L, AD, d, arr, i)
        keep_o, kept_o = oracle_mod.get_L_keep(L, AD, d_o, arr_o, i)
        assert keep.dtype == np.int32 and kept == kept_o and np.array_equal(keep, keep_o)
        fac, like, idx = zs.get_factorials(arr, d, 0.01)
        fac_o, like_o, idx_o = oracle_mod.get_factorials(arr_o, d_o, 0.01)
        assert np.array_equal(fac, fac_o) and np.allclose(like, like_o, rtol=1e-6, atol=1e-9) and np.array_equal(idx, idx_o)
        a = np.ascontiguousarray(af[keep, :][:, k].reshape(-1))
        w_obs, w_l = zs.get_expected_W_l(L, keep, a, AD, arr, fac, like, idx, 1, i)
        var = zs.get_var_W_l(L, keep, a, AD, arr, fac, like, idx, w_l, 1, i)
        w_obs_arr_o, w_l_o, var_o = (np.zeros(kept, np.float32) for _ in range(3))
        kern.expected_W_l(L, keep_o, a, AD, arr_o, fac_o, like_o, idx_o, 1, i, w_obs_arr_o, w_l_o)
        kern.variance_W_l(L, keep_o, a, AD, arr_o, fac_o, like_o, idx_o, 1, i, var_o, w_l_o)
        assert w_l.dtype == np.float32 and w_l.shape == (kept,)
        assert np.max(np.abs(w_l - w_l_o)) <= 2e-6 * np.max(np.abs(w_l_o))
        assert np.max(np.abs(var - var_o)) <= 1e-5 * np.max(np.abs(var_o))
        assert abs(float(w_obs) - float(np.sum(w_obs_arr_o, dtype=np.float32))) <= 1e-6 * abs(float(np.sum(w_obs_arr_o)))
        z = (w_obs - np.sum(w_l)) / np.sqrt(np.sum(var))
        z_ref = float(str(zgold["z_assign_txt"]).split()[i])
        w, mu, v = zgold["z_assign_components"][i][:3]
        assert abs(float(z) - z_ref) <= 2e-6 * (abs(w) + abs(mu)) / np.sqrt(v) + 1e-5 * abs(z_ref)


def test_per_individual_asserts_like_reference(zs, zgold):
    L, AD = zgold["L"], zgold["AD"].astype(np.int32)
    with pytest.raises(AssertionError, match="loci were kept"):
        zs.AD_summary(L, AD, 0, 10 ** 9, False)
