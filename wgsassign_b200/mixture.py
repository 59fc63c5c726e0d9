"""Mixture proportions from an assignment log-likelihood matrix (reference mixture.py).

Host NumPy by design: the input is the tiny [individuals, populations] matrix, not per-SNP
data.  ``em_mix`` keeps the reference's semantics exactly,
including the un-normalised ``exp(loglik)`` (mixture.py:29-30).
"""
import numpy as np


def em_mix(L_mat, L_mat_index, iter, logsumexp=False):
    """Fixed-iteration EM per harvest group (mixture.py:10-39).  Returns a string array
    [groups, 1 + sources]: group name, then the mixing proportions of the last iteration.

    logsumexp=False is the reference: ``exp(loglik)`` un-normalised (mixture.py:29-30), which underflows to 0 for
    genome-scale log-likelihoods (every entry below about -745) and then yields NaN rows.  logsumexp=True
    (``--em_mix_logsumexp``) subtracts each individual's largest log-likelihood first: the assignment probabilities
    ``like * pi / sum(like * pi)`` are invariant under a per-row factor, so wherever the reference's arithmetic is
    finite the two agree to rounding, and the hardened form stays finite everywhere."""
    n_source = L_mat.shape[1]
    groups = np.unique(L_mat_index[:, 1])
    props = np.empty((len(groups), n_source), np.float32)
    for g, name in enumerate(groups):
        ll = np.ascontiguousarray(L_mat[np.flatnonzero(L_mat_index[:, 1] == name), :])
        if logsumexp:
            ll = ll - np.max(ll, axis=1, keepdims=True)
        like = np.exp(ll)
        pi = np.full(n_source, 1.0 / n_source)
        for _ in range(iter):
            post = like * pi                                  # == like @ diag(pi)
            post = post / post.sum(axis=1, keepdims=True)
            pi = post.sum(axis=0) / like.shape[0]
        props[g, :] = pi
    return np.hstack((groups.reshape(-1, 1), props))


def mcmc_mix(L_mat, L_mat_index, iter, seed=None):
    """Gibbs sampler variant (mixture.py:41-77).  The reference's version raises
    UnboundLocalError at its last line (mixture.py:75); this one returns the last draw."""
    n_source = L_mat.shape[1]
    groups = np.unique(L_mat_index[:, 1])
    out = np.empty((len(groups), n_source), np.float32)
    rng = np.random.default_rng(seed)
    for g, name in enumerate(groups):
        like = np.exp(np.ascontiguousarray(L_mat[np.flatnonzero(L_mat_index[:, 1] == name), :]))
        pi = np.full(n_source, 1.0 / n_source)
        for _ in range(iter):
            post = like * pi
            post = post / post.sum(axis=1, keepdims=True)
            counts = rng.multinomial(1, post).sum(axis=0) + 0.001
            pi = rng.dirichlet(counts, 1).reshape(-1)
        out[g, :] = pi
    return np.hstack((groups.reshape(-1, 1), out))
