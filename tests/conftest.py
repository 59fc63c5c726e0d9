import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def bundled():
    return np.load(os.path.join(GOLDEN, "bundled.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def zgold():
    return np.load(os.path.join(GOLDEN, "zscore.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure).  Builds liboracle.so on first use."""
    import subprocess
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    from oracle import oracle
    return oracle


def parse_tsv(text):
    """Parse a LOO tsv (utils.py:113-121 format) -> (header, rows of str)."""
    lines = [l.split("\t") for l in text.strip().split("\n")]
    return lines[0], lines[1:]
