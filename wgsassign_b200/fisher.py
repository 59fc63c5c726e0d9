"""Drop-in for the reference's ``WGSassign/fisher.py`` (fisher_obs :11-43, fisher_obs_ind :45-59).

Both quantities come out of one fused pass over the GL matrix; the second entry point
re-uses the pass of the first when it is called with the same inputs.
"""
import numpy as np

from . import dist, session

_last = {"key": None, "ne_ind": None}


def _run(L, af, IDs):
    pop_of_ind, pops = session.pops_from_ids(IDs)
    ctx = session.context(L, pop_of_ind, len(pops))
    af32 = np.ascontiguousarray(af, dtype=np.float32)
    f_obs, ne_obs, ind_sum = ctx.fisher_partial(af32)
    dist.combine(ctx, ind_sum)
    ne_ind = (ind_sum / float(dist.total_sites(L.shape[0]))).astype(np.float32)
    _last["key"] = (session._sig(L), session._sig(af32), IDs[:, 1].tobytes())
    _last["ne_ind"] = ne_ind
    return f_obs, ne_obs, ne_ind


def fisher_obs(L, af, IDs, t):
    """Observed Fisher information and effective sample size per (site, population):
    two float32 [M, K] arrays."""
    f_obs, ne_obs, _ = _run(L, af, IDs)
    return f_obs, ne_obs


def fisher_obs_ind(L, af, IDs, t):
    """Per-individual effective sample size (mean over sites): float32 [N]."""
    af32 = np.ascontiguousarray(af, dtype=np.float32)
    if _last["key"] == (session._sig(L), session._sig(af32), IDs[:, 1].tobytes()):
        return _last["ne_ind"].copy()
    return _run(L, af, IDs)[2]
