#!/usr/bin/env python
"""How far is the reference from ITSELF in exact arithmetic?  The tolerances of the GPU parity tests that are looser
than 1e-6 (leave-one-out likelihoods of populations of hundreds: 5e-6; Fisher information: 1e-4 of the column scale) are
justified by the rounding noise of the reference's own float32 arithmetic at those shapes.  This script measures it:
the reference's compiled kernels (oracle/_ref; float32 accumulators, double sub-expressions) against a float64
restatement of the same formulas on the same inputs.  Test infrastructure (uses oracle/).
Output: profiles/reference_noise_r2.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from wgsassign_b200 import synth  # noqa: E402


def em64(Lp, iters, tole):
    """emMAF in float64 with the reference's stop rule evaluated in float64."""
    g0, g1 = Lp[:, 0::2].astype(np.float64), Lp[:, 1::2].astype(np.float64)
    g2 = 1.0 - g0 - g1
    f = np.full(Lp.shape[0], 0.25)
    for it in range(iters):
        fp = f.copy()
        p0 = g0 * ((1 - f) ** 2)[:, None]
        p1 = g1 * (2 * f * (1 - f))[:, None]
        p2 = g2 * (f ** 2)[:, None]
        f = np.sum((p1 + 2 * p2) / (2 * (p0 + p1 + p2)), axis=1) / g0.shape[1]
        if np.sqrt(np.mean((f - fp) ** 2)) < tole:
            return f, it + 1
    return f, 0


def loglike64(L, a, i):
    g0, g1 = L[:, 2 * i].astype(np.float64), L[:, 2 * i + 1].astype(np.float64)
    return np.sum(np.log(g0 * (1 - a) ** 2 + g1 * 2 * a * (1 - a) + (1 - g0 - g1) * a ** 2))


def fisher64(Lp, th):
    g0, g1 = Lp[:, 0::2].astype(np.float64), Lp[:, 1::2].astype(np.float64)
    g2 = 1.0 - g0 - g1
    th = th[:, None].astype(np.float64)
    u = g0 * (1 - th) ** 2 + g1 * 2 * th * (1 - th) + g2 * th ** 2
    n1 = 2 * (g0 + g2 - 2 * g1)
    n2 = th * n1 + 2 * (g1 - g0)
    return np.sum(-(n1 / u - (n2 / u) ** 2), axis=1)


def main():
    kern = oracle.kernels("ref" if oracle.have_ref() else "port")
    out = []
    for n_big in (130, 301, 522):
        m, n_small = 120, 10
        d = synth.synth(m, n_big + n_small, 2, seed=51, with_ad=False)
        L, IDs = d["L"], d["IDs"].copy()
        IDs[:n_big, 1] = "big"; IDs[n_big:, 1] = "small"
        af, pops, _ = oracle.reference_af(L, IDs, 200, 1e-4, 4, kern)
        # leave-one-out likelihood of the first 12 members of the big population under their own population, float32 path vs float64
        worst_ll, worst_af = 0.0, 0.0
        for i in range(12):
            cols = [c for j in range(n_big) if j != i for c in (2 * j, 2 * j + 1)]
            Lp = np.ascontiguousarray(L[:, cols])
            f32, it32 = oracle.emMAF(Lp, 200, 1e-4, 4, kern)
            f64, it64 = em64(Lp, 200, 1e-4)
            lo = 1.0 / (2 * n_big)
            a32 = np.clip(f32.astype(np.float64), np.float32(lo), np.float32(1 - lo))
            a64 = np.clip(f64, lo, 1 - lo)
            if it32 == it64:
                worst_af = max(worst_af, float(np.max(np.abs(a32 - a64))))
                ll32, ll64 = loglike64(L, a32, i), loglike64(L, a64, i)
                worst_ll = max(worst_ll, abs(ll32 - ll64) / abs(ll64))
        f_ref, _ = oracle.fisher_obs(L, af, IDs, 4, kern)
        k = list(pops).index("big")
        f64v = fisher64(np.ascontiguousarray(L[:, :2 * n_big]), af[:, k])
        ferr = float(np.max(np.abs(f_ref[:, k] - f64v)) / np.max(np.abs(f64v)))
        out.append("population of %3d, %d sites: reference float32 vs float64 restatement - LOO allele frequency max |diff| %.2e, "
                   "LOO log-likelihood max rel diff %.2e (test tolerance 5e-6), Fisher information max |diff| / column scale %.2e "
                   "(test tolerance 1e-4)" % (n_big, m, worst_af, worst_ll, ferr))
    d = synth.synth(2000, 24, 3, seed=1, with_ad=False)
    af, pops, _ = oracle.reference_af(d["L"], d["IDs"], 200, 1e-4, 4, kern)
    f_ref, _ = oracle.fisher_obs(d["L"], af, d["IDs"], 4, kern)
    pop_of = np.searchsorted(pops, d["IDs"][:, 1])
    worst = 0.0
    for k in range(3):
        cols = [c for j in np.flatnonzero(pop_of == k) for c in (2 * j, 2 * j + 1)]
        f64v = fisher64(np.ascontiguousarray(d["L"][:, cols]), af[:, k])
        worst = max(worst, float(np.max(np.abs(f_ref[:, k] - f64v)) / np.max(np.abs(f64v))))
    out.append("populations of 8, 2000 sites: Fisher information reference float32 vs float64: max |diff| / column scale %.2e" % worst)
    text = "\n".join(out)
    print(text)
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "reference_noise_r2.txt"), "w") as fh:
        fh.write("scripts/reference_noise.py: rounding noise of the reference's own float32 arithmetic (its compiled kernels vs a float64 restatement)\n" + text + "\n")


if __name__ == "__main__":
    main()
