"""wgsassign_b200 - B200-native implementation of WGSassign's genotype-likelihood hot path.

Same entry points as the reference package (``emMAF``, ``glassy``, ``fisher``, ``zscore``,
``mixture``, ``utils``, ``reader`` and the ``WGSassign`` CLI); the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI declared in ``include/wgsassign_b200.h``.
"""
__all__ = ["emMAF", "glassy", "fisher", "session", "dist", "synth"]
