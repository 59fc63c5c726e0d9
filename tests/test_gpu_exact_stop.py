"""The EM stop rule at scale: the reference adds the squared changes of all sites into ONE float32 accumulator,
left to right (rmse1d, emMAF_cy.pyx:26-33).  At millions of sites that sum is biased against the exact one (-0.4 %
at 5 M sites), enough to move the stop iteration, so checks that land inside the band where the exact
FP64 sum cannot decide are resolved on the device with an order-exact emulation of that loop (block_seqsum32).

  * the primitive against a serial float32 loop, bit for bit, on adversarial vectors;
  * `--get_reference_af` + `--loo` on 6 M device-generated sites against the oracle run on the downloaded matrix:
    every stop iteration equal, allele frequencies within 1e-5.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import wgsassign_b200._lib as _lib
    if _lib.lib().wgs_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    return _lib


def serial_sum(x, carry=0.0):
    """float32 left-to-right: numpy's cumsum on a float32 array adds sequentially in float32."""
    v = np.concatenate(([np.float32(carry)], np.asarray(x, np.float32)))
    return np.cumsum(v, dtype=np.float32)[-1]


def test_serial_sum_helper_is_sequential():
    rng = np.random.default_rng(0)
    x = (rng.random(3000) * 1e-3).astype(np.float32)
    r = np.float32(0)
    for v in x:
        r = np.float32(r + v)
    assert serial_sum(x).tobytes() == r.tobytes()


@pytest.mark.parametrize("kind", ["squares", "mixed_magnitudes", "constant", "ties", "zeros", "huge_n", "specials"])
def test_seqsum_matches_serial_loop(lib, kind):
    rng = np.random.default_rng(sum(map(ord, kind)))
    ctx = lib.Context(0)
    cases = []
    if kind == "squares":                         # what the stop rule adds: squared AF changes of ~1e-4
        cases = [((rng.standard_normal(n) * 1e-4) ** 2).astype(np.float32) for n in (1, 31, 256, 257, 100003)]
    elif kind == "mixed_magnitudes":
        cases = [(rng.random(n) * rng.choice([1e-12, 1e-3, 1.0, 1e5], n)).astype(np.float32) for n in (777, 40000)]
    elif kind == "constant":                      # stagnation: the addend falls below half an ulp of the sum
        cases = [np.full(n, np.float32(0.333333), np.float32) for n in (5000, 3000000)]
    elif kind == "ties":                          # exact half-way cases: round-half-even depends on the running parity
        cases = [(rng.integers(0, 3, n) * np.float32(0.5) + rng.integers(0, 2, n) * np.float32(2.0 ** -20)).astype(np.float32)
                 for n in (4096, 50001)]
    elif kind == "zeros":
        x = (rng.random(9000) * 1e-6).astype(np.float32)
        x[rng.integers(0, 9000, 4000)] = 0
        cases = [x, np.zeros(1000, np.float32), np.zeros(0, np.float32)]
    elif kind == "huge_n":                        # 20 M addends: the reference's own sum is 2 % off the exact one here
        cases = [((rng.standard_normal(20_000_000) * 1e-4) ** 2).astype(np.float32)]
    else:                                         # negative, infinite and NaN addends take the in-order path
        x = (rng.random(3000) * 1e-3).astype(np.float32)
        x[100] = -1e-4
        y = x.copy(); y[2000] = np.inf
        z = x.copy(); z[17] = np.nan
        cases = [x, y, z]
    for x in cases:
        for carry in (0.0, 1e-3):
            got, ref = ctx.debug_seqsum(x, carry), serial_sum(x, carry)
            assert np.float32(got).tobytes() == np.float32(ref).tobytes() or (np.isnan(got) and np.isnan(ref)), (kind, len(x), carry, got, ref)
    if kind == "huge_n":
        x = cases[0]
        exact = float(np.sum(x, dtype=np.float64))
        assert abs(float(serial_sum(x)) / exact - 1.0) > 1e-3        # the bias this machinery exists for
    ctx.close()


def test_stop_iterations_match_oracle_at_6m_sites(lib, oracle_mod):
    m, n, k = 6_000_000, 32, 2
    pop_of = ((np.arange(n) * k) // n).astype(np.int32)
    ids = np.empty((n, 2), dtype="U16")
    for i in range(n):
        ids[i, 0], ids[i, 1] = "ind%d" % i, "pop%02d" % pop_of[i]
    ctx = lib.Context(0)
    ctx.set_pops(pop_of, k)
    ctx.synth(m, n, seed=99)
    af, its = ctx.ref_af(200, 1e-4)
    a_gpu = af.copy()
    ll, _, lits = ctx.loo_partial(a_gpu, 200, 1e-4)
    # every check resolved sequentially (-1), and the rigorous summation bound as the band (-2): the default band
    # (8x the measured float32 bias) must not change a single decision
    for band in (-1, -2):
        ctx.set_option("rmse_band_ppm", band)
        af_all, its_all = ctx.ref_af(200, 1e-4)
        a_all = af_all.copy()
        ll_all, _, lits_all = ctx.loo_partial(a_all, 200, 1e-4)
        assert list(its_all) == list(its) and list(lits_all) == list(lits) and np.array_equal(af_all, af) and np.array_equal(a_all, a_gpu)
    # the FP64-only rule, for the record (it may or may not differ on this seed)
    ctx.set_option("rmse_band_ppm", 0)
    ctx.set_option("rmse_exact", 0)
    _, its64 = ctx.ref_af(200, 1e-4)
    _, _, lits64 = ctx.loo_partial(af.copy(), 200, 1e-4)
    ctx.set_option("rmse_exact", 1)
    L = np.empty((m, 2 * n), np.float32)
    for s0 in range(0, m, 500_000):
        L[s0:s0 + 500_000] = ctx.download(s0, min(500_000, m - s0))
    ctx.close()
    import os
    t = os.cpu_count() or 1
    af_o, _, its_o = oracle_mod.reference_af(L, ids, 200, 1e-4, t)
    a_o = af_o.copy()
    ll_o, _, lits_o = oracle_mod.loo(L, a_o, ids, t, 200, 1e-4)
    print("stop iterations  oracle:", list(its_o), list(lits_o))
    print("            FP64 sums :", list(its64), list(lits64))
    assert list(its) == list(its_o)
    assert list(lits) == list(lits_o)
    assert np.max(np.abs(af - af_o)) < 1e-5 and np.max(np.abs(a_gpu - a_o)) < 1e-5
    assert np.max(np.abs(ll - ll_o) / np.abs(ll_o)) < 1e-6 and np.array_equal(np.argmax(ll, 1), np.argmax(ll_o, 1))
