"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI via the
drop-in entry points, against the golden fixtures of the unmodified reference and against
the CPU oracle on seeded synthetic inputs.

Tolerances are BASELINE.json's: assignment argmax / LOO assignments / EM iteration counts
bit-exact; allele frequencies within 1e-5 absolute; log-likelihoods within 1e-6 relative.
"""
import numpy as np
import pytest

from conftest import parse_tsv

pytestmark = pytest.mark.gpu

AF_ATOL = 1e-5
LL_RTOL = 1e-6
FISHER_TOL = 2e-5       # relative to |value| + 1e-3 of the column scale; the reference itself: 0.7e-6 .. 5.4e-6 (see test_config1)


@pytest.fixture(scope="module")
def wgs():
    import wgsassign_b200._lib as _lib
    if _lib.lib().wgs_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    from wgsassign_b200 import emMAF, fisher, glassy, session
    session.reset()

    class NS:
        pass
    ns = NS()
    ns.lib, ns.emMAF, ns.fisher, ns.glassy, ns.session = _lib, emMAF, fisher, glassy, session
    yield ns
    session.reset()


def rel_err(a, b):
    return np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.abs(b.astype(np.float64)), 1e-30))


def test_repack_roundtrip(wgs):
    from wgsassign_b200 import synth
    d = synth.synth(301, 23, 4, seed=3, interleave=True)
    L, AD = d["L"], d["AD"]
    pop_of, pops = wgs.session.pops_from_ids(d["IDs"])
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl(L)
    ctx.upload_ad(AD)
    L2, AD2 = ctx.download(0, 301, want_ad=True)
    assert np.array_equal(L, L2) and np.array_equal(AD, AD2)
    L3 = ctx.download(100, 50)
    assert np.array_equal(L[100:150], L3)
    ctx.close()


def test_config2_pop_like_bundled(wgs, bundled):
    ll = wgs.glassy.assignLL(bundled["L_nonbreeding"], bundled["c1_pop_af"], 1)
    gold = bundled["c2_pop_like"]
    assert ll.dtype == np.float32 and ll.shape == (34, 5)
    assert rel_err(ll, gold) < LL_RTOL
    assert "".join(str(int(x)) for x in np.argmax(ll, 1)) == "1120222122011010110000333013331333"


def test_config1_reference_af_bundled(wgs, bundled, capsys):
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    pop_of, pops = wgs.session.pops_from_ids(IDs)
    ctx = wgs.session.context(L, pop_of, len(pops))
    af, its = ctx.ref_af(200, 1e-4)
    assert list(its) == list(bundled["c1_em_iters_ref"])
    assert np.max(np.abs(af - bundled["c1_pop_af"])) < AF_ATOL
    assert af.min() == np.float32(1 / 48) and af.max() == bundled["c1_pop_af"].max()
    # the emMAF drop-in on one population's column copy (emMAF.py:15)
    cols = np.flatnonzero(np.repeat(IDs[:, 1] == pops[0], 2))
    f = wgs.emMAF.emMAF(np.ascontiguousarray(L[:, cols]), 200, 1e-4, 1)
    assert "EM (MAF) converged at iteration: 17" in capsys.readouterr().out
    n0 = int(np.sum(IDs[:, 1] == pops[0]))
    clipped = np.clip(f, np.float32(1 / (2 * (n0 + 1))), np.float32(1 - 1 / (2 * (n0 + 1))))
    assert np.max(np.abs(clipped - bundled["c1_pop_af"][:, 0])) < AF_ATOL


def test_config1_loo_bundled(wgs, bundled):
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    af = bundled["c1_pop_af"].copy()
    ll, parts = wgs.glassy.loo(L, af, IDs, 1, 200, 1e-4)
    hdr, rows = parse_tsv(str(bundled["c1_loo_tsv"]))
    gold = np.array([[float(x) for x in r[2:]] for r in rows])
    assert rel_err(ll, gold) < 2e-6          # the golden TSV itself carries only 6 decimals
    assert np.array_equal(np.argmax(ll, 1), np.argmax(gold, 1))
    assert "".join(str(int(x)) for x in np.argmax(ll, 1)) == \
        "2222222422333333333333333311111111113333333444401010010100000111111111144444444422222"
    assert np.array_equal(parts, ll)


def test_config1_loo_matches_oracle_exact_quantities(wgs, bundled, oracle_mod):
    """LOO against the oracle directly: iteration counts, full-precision likelihoods and
    the mutated AF matrix the reference leaves behind (glassy.py:89)."""
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    af_o = bundled["c1_pop_af"].copy()
    ll_o, _, its_o = oracle_mod.loo(L, af_o, IDs, 1, 200, 1e-4)
    pop_of, pops = wgs.session.pops_from_ids(IDs)
    ctx = wgs.session.context(L, pop_of, len(pops))
    af_g = bundled["c1_pop_af"].copy()
    ll_g, _, its_g = ctx.loo_partial(af_g, 200, 1e-4)
    assert list(its_g) == list(its_o) == list(bundled["c1_em_iters_loo"])
    assert rel_err(ll_g, ll_o) < LL_RTOL
    assert np.max(np.abs(af_g - af_o)) < AF_ATOL


def test_config1_downsampled_partitions(wgs, bundled):
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    common = np.isin(bundled["sites_breeding"], bundled["sites_breeding_ds"])
    Lf = np.ascontiguousarray(L[common])
    af = bundled["c1ds_pop_af"].copy()
    ll, parts = wgs.glassy.loo(Lf, af, IDs, 1, 200, 1e-4, downsampled_L=bundled["L_breeding_ds"], num_partitions=3)
    _, rows = parse_tsv(str(bundled["c1ds_loo_tsv"]))
    gold = np.array([[float(x) for x in r[2:]] for r in rows])
    assert rel_err(ll, gold) < 2e-6
    assert np.array_equal(np.argmax(ll, 1), np.argmax(gold, 1))
    _, prow = parse_tsv(str(bundled["c1ds_loo_parts_tsv"]))
    pgold = np.array([[float(x) for x in r[3:]] for r in prow])
    assert parts.shape == pgold.shape == (85 * 3, 5)
    assert rel_err(parts, pgold) < 5e-6
    # size-independent property: the modulo partitions tile the sites exactly once
    assert rel_err(parts.reshape(85, 3, 5).astype(np.float64).sum(1), ll) < 1e-6


def test_config1_fisher_bundled(wgs, bundled):
    L, IDs, af = bundled["L_breeding"], bundled["IDs_breeding"], bundled["c1_pop_af"]
    f_obs, ne_obs = wgs.fisher.fisher_obs(L, af, IDs, 1)
    ne_ind = wgs.fisher.fisher_obs_ind(L, af, IDs, 1)
    gf, gn = bundled["c1_fisher_obs"], bundled["c1_ne_obs"]
    scale_f, scale_n = np.abs(gf).max(0), np.abs(gn).max(0)
    # FISHER_TOL: the reference's own float32 result is 0.7e-6 .. 5.4e-6 away from a float64 restatement under this metric
    # (populations of 8 .. 522; scripts/reference_noise.py, profiles/reference_noise_r2.txt)
    assert np.max(np.abs(f_obs - gf) / (np.abs(gf) + 1e-3 * scale_f)) < FISHER_TOL
    assert np.max(np.abs(ne_obs - gn) / (np.abs(gn) + 1e-3 * scale_n)) < FISHER_TOL
    assert rel_err(ne_obs.mean(0), gn.mean(0)) < 1e-5
    gold_ind = np.array([float(x) for x in str(bundled["c1_ne_ind_txt"]).split()])
    assert rel_err(ne_ind, gold_ind) < 1e-5


@pytest.mark.parametrize("m,n,k,interleave,seed", [
    (1000, 37, 3, True, 11),        # ragged: N not a multiple of 4/32, interleaved populations
    (2500, 130, 20, False, 12),     # K = 20 tile, several column groups
    (777, 9, 2, True, 13),          # tiny populations
    (64, 70, 7, False, 14),         # fewer sites than one tile
])
def test_synthetic_vs_oracle(wgs, oracle_mod, m, n, k, interleave, seed):
    from wgsassign_b200 import synth
    d = synth.synth(m, n, k, seed=seed, interleave=interleave, with_ad=False)
    L, IDs = d["L"], d["IDs"]
    af_o, pops, its_o = oracle_mod.reference_af(L, IDs, 200, 1e-4, 2)
    pop_of, pops2 = wgs.session.pops_from_ids(IDs)
    ctx = wgs.session.context(L, pop_of, len(pops2))
    af_g, its_g = ctx.ref_af(200, 1e-4)
    assert list(its_g) == list(its_o)
    assert np.max(np.abs(af_g - af_o)) < AF_ATOL
    # pop_like on the oracle's AF
    ll_o = oracle_mod.assignLL(L, af_o, 2)
    ll_g = wgs.glassy.assignLL(L, af_o, 1)
    assert rel_err(ll_g, ll_o) < LL_RTOL
    assert np.array_equal(np.argmax(ll_g, 1), np.argmax(ll_o, 1))
    # fisher
    f_o, ne_o = oracle_mod.fisher_obs(L, af_o, IDs, 2)
    f_g, ne_g = wgs.fisher.fisher_obs(L, af_o, IDs, 1)
    assert np.max(np.abs(f_g - f_o) / (np.abs(f_o) + 1e-3 * np.abs(f_o).max(0))) < FISHER_TOL
    assert rel_err(wgs.fisher.fisher_obs_ind(L, af_o, IDs, 1), oracle_mod.fisher_obs_ind(L, af_o, IDs, 2)) < 1e-5
    # LOO
    if n <= 40:
        a1, a2 = af_o.copy(), af_o.copy()
        llo, _, itso = oracle_mod.loo(L, a1, IDs, 2, 200, 1e-4)
        llg, _, itsg = ctx.loo_partial(a2, 200, 1e-4)
        assert list(itsg) == list(itso)
        assert rel_err(llg, llo) < LL_RTOL
        assert np.array_equal(np.argmax(llg, 1), np.argmax(llo, 1))
        assert np.max(np.abs(a1 - a2)) < AF_ATOL


def test_edge_cases(wgs, oracle_mod):
    """Boundary allele frequencies (log 0 = -inf like the reference), missing data rows,
    a single-member population (0/0 = NaN in the reference's LOO) and an empty site set."""
    from wgsassign_b200 import synth
    d = synth.synth(200, 10, 2, seed=21, with_ad=False)
    L, IDs = d["L"], d["IDs"]
    A = np.full((200, 2), 0.3, np.float32)
    A[5, 0] = 0.0
    A[9, 1] = 1.0
    L2 = L.copy()
    L2[5, 0:2] = (0.0, 0.0)       # individual 0 is hom-alt at a site where pop 0 has a = 0
    L2[:, 6:8] = 0.333333         # individual 3 entirely missing
    ll_o = oracle_mod.assignLL(L2, A, 1)
    ll_g = wgs.glassy.assignLL(L2, A, 1)
    assert np.isneginf(ll_o[0, 0]) and np.isneginf(ll_g[0, 0])
    fin = np.isfinite(ll_o)
    assert np.array_equal(fin, np.isfinite(ll_g))
    assert rel_err(ll_g[fin], ll_o[fin]) < LL_RTOL
    # single-member population
    IDs1 = IDs.copy()
    IDs1[:, 1] = "a"
    IDs1[9, 1] = "b"
    af_o, _, _ = oracle_mod.reference_af(L, IDs1, 200, 1e-4, 1)
    a1, a2 = af_o.copy(), af_o.copy()
    llo, _, itso = oracle_mod.loo(L, a1, IDs1, 1, 20, 1e-4)
    pop_of, pops = wgs.session.pops_from_ids(IDs1)
    ctx = wgs.session.context(L, pop_of, len(pops))
    llg, _, itsg = ctx.loo_partial(a2, 20, 1e-4)
    assert list(itsg) == list(itso)
    assert np.array_equal(np.isnan(llo), np.isnan(llg))
    ok = ~np.isnan(llo)
    assert rel_err(llg[ok], llo[ok]) < LL_RTOL


def test_wrong_dtype_rejected(wgs):
    with pytest.raises(ValueError):
        wgs.glassy.assignLL(np.zeros((4, 4), np.float64), np.zeros((4, 1), np.float32), 1)
    with pytest.raises(ValueError):
        wgs.glassy.assignLL(np.zeros((8, 4), np.float32)[::2], np.zeros((4, 1), np.float32), 1)


def test_determinism_and_partition_property_large(wgs):
    """Size-independent properties on a device-generated matrix larger than L2:
    bitwise run-to-run determinism, and modulo partitions summing to the whole."""
    ctx = wgs.lib.Context(0)
    n, k, m = 500, 10, 200_000
    pop_of = ((np.arange(n) * k) // n).astype(np.int32)
    ctx.set_pops(pop_of, k)
    ctx.synth(m, n, seed=7)
    af, its = ctx.ref_af(200, 1e-4)
    assert af.min() >= np.float32(1 / 102) and af.max() <= np.float32(1 - 1 / 102)
    ll1 = ctx.pop_like_partial(af)
    ll2 = ctx.pop_like_partial(af)
    assert np.array_equal(ll1, ll2)
    # individuals are most likely under their own population on depth-consistent data
    assert np.mean(np.argmax(ll1, 1) == pop_of) > 0.99
    ctx.close()


@pytest.mark.parametrize("interleave,pinned", [(False, True), (False, False), (True, True)])
def test_async_upload_roundtrip(wgs, interleave, pinned):
    """wgs_upload_gl_async (one strided DMA per population slab, straight into the device layout) leaves
    exactly the matrix wgs_upload_gl leaves: contiguous populations (slab path), pageable memory, and an
    interleaved ID file (falls back to the chunked upload)."""
    from wgsassign_b200 import synth
    d = synth.synth(1201, 23, 4, seed=5, interleave=interleave, with_ad=False)
    L = d["L"]
    if pinned:
        Lp = wgs.lib.pinned_empty(L.shape, np.float32)
        Lp[...] = L
    else:
        Lp = L
    pop_of, pops = wgs.session.pops_from_ids(d["IDs"])
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl_async(Lp)
    L2 = ctx.download(0, 1201)               # every operator but ref_af_loo waits for the whole matrix
    assert np.array_equal(L, L2)
    ctx.upload_gl_async(Lp)
    ctx.upload_wait()
    assert np.array_equal(L[7:19], ctx.download(7, 12))
    ctx.close()


@pytest.mark.parametrize("m,n,k,interleave,parts,seed", [
    (3000, 37, 3, False, 1, 31),
    (1500, 64, 5, False, 3, 32),
    (900, 21, 4, True, 1, 33),
])
def test_ref_af_loo_fused_is_bit_identical(wgs, m, n, k, interleave, parts, seed):
    """wgs_ref_af_loo == wgs_ref_af followed by wgs_loo_partial, bit for bit, whether the matrix was
    uploaded synchronously or is still arriving slab by slab (pipelined leave-one-out EM)."""
    from wgsassign_b200 import synth
    d = synth.synth(m, n, k, seed=seed, interleave=interleave, with_ad=False)
    L = wgs.lib.pinned_empty(d["L"].shape, np.float32)
    L[...] = d["L"]
    pop_of, pops = wgs.session.pops_from_ids(d["IDs"])
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl(L)
    af0, its0 = ctx.ref_af(200, 1e-4)
    af_mut = af0.copy()
    ll0, llp0, lits0 = ctx.loo_partial(af_mut, 200, 1e-4, parts=parts)
    for use_async in (False, True):
        ctx.set_pops(pop_of, len(pops))
        if use_async:
            ctx.upload_gl_async(L)
        else:
            ctx.upload_gl(L)
        af1, its1, ll1, llp1, lits1, af_after = ctx.ref_af_loo(200, 1e-4, parts=parts, want_af_after=True)
        assert np.array_equal(af1, af0) and list(its1) == list(its0)
        assert np.array_equal(ll1, ll0) and np.array_equal(llp1, llp0)
        assert list(lits1) == list(lits0)
        assert np.array_equal(af_after, af_mut)
    ctx.close()


def test_ref_af_loo_fused_bundled(wgs, bundled):
    """The fused call on the bundled breeding data against the reference's golden outputs."""
    L, IDs = bundled["L_breeding"], bundled["IDs_breeding"]
    Lp = wgs.lib.pinned_empty(L.shape, np.float32)
    Lp[...] = L
    pop_of, pops = wgs.session.pops_from_ids(IDs)
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl_async(Lp)
    af, its, ll, _, lits, _ = ctx.ref_af_loo(200, 1e-4)
    assert list(its) == list(bundled["c1_em_iters_ref"]) and list(lits) == list(bundled["c1_em_iters_loo"])
    assert np.max(np.abs(af - bundled["c1_pop_af"])) < AF_ATOL
    _, rows = parse_tsv(str(bundled["c1_loo_tsv"]))
    gold = np.array([[float(x) for x in r[2:]] for r in rows])
    assert rel_err(ll.astype(np.float32), gold) < 2e-6
    assert np.array_equal(np.argmax(ll, 1), np.argmax(gold, 1))
    ctx.close()


def test_fused_edge_cases(wgs, oracle_mod):
    """Pipelined path on awkward shapes: a single-member population (NaN columns like the reference's 0/0), one
    population only, and fewer sites than one tile."""
    from wgsassign_b200 import synth
    d = synth.synth(150, 11, 3, seed=41, with_ad=False)
    L = wgs.lib.pinned_empty(d["L"].shape, np.float32)
    L[...] = d["L"]
    IDs = d["IDs"].copy()
    IDs[:, 1] = "a"
    IDs[4:8, 1] = "b"
    IDs[10, 1] = "c"                                  # single member
    for ids in (IDs, np.array([[x, "p"] for x in IDs[:, 0]])):
        pop_of, pops = wgs.session.pops_from_ids(ids)
        ctx = wgs.lib.Context(0)
        ctx.set_pops(pop_of, len(pops))
        ctx.upload_gl_async(L)
        af, its, ll, _, lits, af_after = ctx.ref_af_loo(30, 1e-4, want_af_after=True)
        af_o, _, its_o = oracle_mod.reference_af(d["L"], ids, 30, 1e-4, 1)
        a1 = af_o.copy()
        ll_o, _, lits_o = oracle_mod.loo(d["L"], a1, ids, 1, 30, 1e-4)
        assert list(its) == list(its_o) and list(lits) == list(lits_o)
        assert np.max(np.abs(af - af_o)) < AF_ATOL
        assert np.array_equal(np.isnan(ll), np.isnan(ll_o))
        ok = ~np.isnan(ll_o)
        assert rel_err(ll[ok], ll_o[ok]) < LL_RTOL
        fin = ~np.isnan(a1)
        assert np.array_equal(np.isnan(af_after), np.isnan(a1)) and np.max(np.abs(af_after[fin] - a1[fin])) < AF_ATOL
        ctx.close()


def test_fallback_kernels_agree(wgs, monkeypatch):
    """The kernels kept for shapes the fast ones do not cover (gather-through-L1 loo_like, TMA-tile Fisher,
    one-iteration-per-launch EM) give the same answers as the fast ones on a shape both cover."""
    from wgsassign_b200 import synth
    d = synth.synth(2100, 45, 4, seed=43, with_ad=False)
    pop_of, pops = wgs.session.pops_from_ids(d["IDs"])
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl(d["L"])
    af, its = ctx.ref_af(200, 1e-4)
    a0 = af.copy()
    ll, _, lits = ctx.loo_partial(a0, 200, 1e-4)
    f_obs, ne_obs, ind = ctx.fisher_partial(af)
    for name in ("loolike_v1", "fisher_v1", "em_step", "em_no_lookahead", "loo_v4", "loo_nofirst"):
        ctx.set_option(name, 1)
    af2, its2 = ctx.ref_af(200, 1e-4)
    a1 = af2.copy()
    ll2, _, lits2 = ctx.loo_partial(a1, 200, 1e-4)
    f2, ne2, ind2 = ctx.fisher_partial(af)
    assert list(its) == list(its2) and list(lits) == list(lits2)
    assert np.max(np.abs(af - af2)) < 1e-6 and np.max(np.abs(a0 - a1)) < AF_ATOL
    assert rel_err(ll2, ll) < LL_RTOL and np.array_equal(np.argmax(ll, 1), np.argmax(ll2, 1))
    assert np.max(np.abs(f2 - f_obs) / (np.abs(f_obs) + 1e-3 * np.abs(f_obs).max(0))) < FISHER_TOL
    assert rel_err(ind2, ind) < 1e-5
    ctx.close()


def test_loo_population_by_population_is_bitwise_identical(wgs):
    """One packed-row buffer, populations one after the other (what a shard too large for all packed rows at once
    runs): the same updates and decisions, hence the same bits, as all populations per iteration."""
    from wgsassign_b200 import synth
    d = synth.synth(3100, 52, 4, seed=47, with_ad=False)
    pop_of, pops = wgs.session.pops_from_ids(d["IDs"])
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl(d["L"])
    af, _ = ctx.ref_af(200, 1e-4)
    a0, a1 = af.copy(), af.copy()
    ll0, _, its0 = ctx.loo_partial(a0, 200, 1e-4)
    ctx.set_option("loo_by_pop", 1)
    ll1, _, its1 = ctx.loo_partial(a1, 200, 1e-4)
    assert list(its0) == list(its1) and np.array_equal(a0, a1) and np.array_equal(ll0, ll1)
    ctx.close()


def test_loo_like_block_shapes_agree(wgs, monkeypatch):
    """The three launch shapes of the staged leave-one-out likelihood kernel (two small blocks per SM, big blocks of up
    to 18 warps, the gather-through-L1 fallback) on a panel wide enough (22 warps of individuals) to split big blocks."""
    ctx = wgs.lib.Context(0)
    n, k, m = 700, 4, 3000
    pop_of = ((np.arange(n) * k) // n).astype(np.int32)
    ctx.set_pops(pop_of, k)
    ctx.synth(m, n, seed=11)
    af, _ = ctx.ref_af(50, 1e-4)
    ll0, _, its0 = ctx.loo_partial(af.copy(), 50, 1e-4)
    assert np.mean(np.argmax(ll0, 1) == pop_of) > 0.99
    ctx.set_option("loolike_v2", 1)                     # the staged state-row kernel itself (population-contiguous panels take loo_like3)
    for name in ("loolike_smallblock", "loolike_v1"):
        ctx.set_option(name, 1)
        ll1, _, its1 = ctx.loo_partial(af.copy(), 50, 1e-4)
        assert list(its1) == list(its0)
        assert rel_err(ll1, ll0) < 1e-7 and np.array_equal(np.argmax(ll1, 1), np.argmax(ll0, 1))
    ctx.close()


@pytest.mark.parametrize("n_big", [522, 301, 130])
def test_large_population_paths_vs_oracle(wgs, oracle_mod, n_big):
    """A population of more than 512 individuals takes the kernels kept for that size (TMA-tile population EM with
    replay, 512-thread leave-one-out blocks with a general first iteration, TMA-tile Fisher pass); 301 and 130 take
    the widest instantiations (32 and 16 threads per site row) of the register-tile kernels."""
    from wgsassign_b200 import synth
    n_small, m = 10, 120
    d = synth.synth(m, n_big + n_small, 2, seed=51, with_ad=False)
    L, IDs = d["L"], d["IDs"].copy()
    IDs[:n_big, 1] = "big"
    IDs[n_big:, 1] = "small"
    af_o, _, its_o = oracle_mod.reference_af(L, IDs, 200, 1e-4, 4)
    pop_of, pops = wgs.session.pops_from_ids(IDs)
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl(L)
    af_g, its_g = ctx.ref_af(200, 1e-4)
    assert list(its_g) == list(its_o)
    assert np.max(np.abs(af_g - af_o)) < AF_ATOL
    a1, a2 = af_o.copy(), af_o.copy()
    ll_o, _, lits_o = oracle_mod.loo(L, a1, IDs, 4, 200, 1e-4)
    ll_g, _, lits_g = ctx.loo_partial(a2, 200, 1e-4)
    assert list(lits_g) == list(lits_o)
    # The reference adds the n posterior terms of a site sequentially in float32 (emMAF_cy.pyx:19-23).  Measured
    # (scripts/reference_noise.py, profiles/reference_noise_r2.txt): against a float64 restatement of its own formulas its
    # leave-one-out frequencies are off by 1.0e-6 / 2.8e-6 / 6.3e-6 at n = 130 / 301 / 522 and its own-population
    # log-likelihoods by 1.3e-7 / 4.6e-7 / 5.6e-7 relative; a frequency near the clip bound 1/(2n) passes its error on to the
    # other populations' likelihoods undamped.  5e-6 here (about 9x the reference's own noise), 1e-6 everywhere else.
    assert rel_err(ll_g, ll_o) < 5e-6 and np.array_equal(np.argmax(ll_g, 1), np.argmax(ll_o, 1))
    assert np.max(np.abs(a1 - a2)) < AF_ATOL
    f_o, _ = oracle_mod.fisher_obs(L, af_o, IDs, 4)
    f_g, _, ind_g = ctx.fisher_partial(af_o)
    assert np.max(np.abs(f_g - f_o) / (np.abs(f_o) + 1e-3 * np.abs(f_o).max(0))) < FISHER_TOL
    ctx.close()


@pytest.mark.parametrize("m,n,k", [(1, 6, 2), (5, 7, 3), (33, 4, 2)])
def test_tiny_shapes_fused_vs_oracle(wgs, oracle_mod, m, n, k):
    """Degenerate sizes (a single site, populations of two, fewer sites than any tile) through the fused call."""
    from wgsassign_b200 import synth
    d = synth.synth(m, n, k, seed=61 + m, with_ad=False)
    L, IDs = d["L"], d["IDs"]
    pop_of, pops = wgs.session.pops_from_ids(IDs)
    ctx = wgs.lib.Context(0)
    ctx.set_pops(pop_of, len(pops))
    ctx.upload_gl_async(L)
    af, its, ll, _, lits, _ = ctx.ref_af_loo(40, 1e-4)
    af_o, _, its_o = oracle_mod.reference_af(L, IDs, 40, 1e-4, 1)
    a1 = af_o.copy()
    ll_o, _, lits_o = oracle_mod.loo(L, a1, IDs, 1, 40, 1e-4)
    assert list(its) == list(its_o) and list(lits) == list(lits_o)
    assert np.max(np.abs(af - af_o)) < AF_ATOL
    ok = np.isfinite(ll_o)
    assert np.array_equal(ok, np.isfinite(ll))
    assert rel_err(ll[ok], ll_o[ok]) < LL_RTOL
    ctx.close()
