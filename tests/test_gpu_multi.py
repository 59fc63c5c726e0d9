"""Site-sharded multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharding_matches_single_gpu():
    import wgsassign_b200._lib as _lib
    n = _lib.lib().wgs_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633", os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "MULTI_GPU_CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
