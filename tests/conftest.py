import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Make sure the native library and the C oracle are built (no-ops when up to date; nvcc/gcc cross-compile on a
    CPU-only box).  A failure here is reported by the tests that need the artefacts, not swallowed."""
    try:
        from liftreg_b200 import build as native_build
        native_build.build()
    except Exception as e:  # pragma: no cover
        print("WARNING: could not build libliftreg_b200.so: %r" % (e,))
    try:
        from oracle import c_oracle
        c_oracle.build()
    except Exception as e:  # pragma: no cover
        print("WARNING: could not build the C oracle: %r" % (e,))
    try:        # the reference's own hot-path modules for the GPU box (git-ignored copy; no-op without /root/reference)
        from oracle import vendor_reference
        vendor_reference.vendor()
    except Exception as e:  # pragma: no cover
        print("WARNING: could not vendor the reference modules: %r" % (e,))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.sqrt((b ** 2).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / den) if den > 0 else float(np.sqrt(((a - b) ** 2).sum()))


@pytest.fixture(scope="session")
def golden():
    return load_golden
