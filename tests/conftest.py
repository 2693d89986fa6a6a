import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "unsupervised-asr_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "eodm_golden.npz"))


@pytest.fixture(scope="session")
def eodm():
    """The product package; importing it loads libeodm_b200.so (built on demand here)."""
    so = os.path.join(ROOT, "unsupervised-asr_b200", "eodm_b200", "libeodm_b200.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "unsupervised-asr_b200", "csrc"), "-j8"])
    import eodm_b200
    return eodm_b200
