import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """The reference's own fixtures (brisk_verification_{ast,harris}.set), see tools/make_golden.py."""
    return np.load(ROOT / "tests" / "golden" / "brisk_verification.npz")


@pytest.fixture(scope="session")
def golden_provided():
    """Outputs of the compiled reference for ComputeScale / passed key points, see tests/golden/make_provided_keypoints.py."""
    return np.load(ROOT / "tests" / "golden" / "provided_keypoints.npz")


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference (oracle/brisk_oracle.cc), built on demand."""
    from oracle import restate
    restate.lib()
    return restate


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (prebuilt; skipped when absent)."""
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libbrisk_ref.so not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def ctx():
    import ethzasl_brisk_b200 as bb
    return bb.Context(0)


def kp_equal(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)
