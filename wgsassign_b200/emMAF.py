"""Drop-in for the reference's ``WGSassign/emMAF.py`` (emMAF.py:15-27)."""
import numpy as np

from . import _lib, session


def emMAF(L, iter, tole, t):
    """EM allele-frequency estimate of the individuals in ``L`` ([M, 2n] float32).

    Same contract as the reference: uniform 0.25 start, at most ``iter`` updates, stop when
    the RMSE between successive estimates over all sites drops below ``tole`` (checked
    after the update, emMAF.py:21-25), prints the converging iteration, returns float32
    [M].  ``t`` (threads) is accepted and ignored: the update runs on the GPU.
    """
    L = np.ascontiguousarray(L)
    ctx = session._state["ctx"] or _lib.Context(session.device_index())
    if session._state["ctx"] is None:
        session._state["ctx"] = ctx
    f, it = ctx.emMAF(L, iter, tole)
    if it > 0:
        print("EM (MAF) converged at iteration: " + str(it))
    return f
