"""pytest configuration: markers, path setup and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def goldens():
    """Literal golden arrays of the reference's own unit tests (tests/golden/extract_reference_goldens.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_goldens.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.lib()
    return orc


@pytest.fixture(scope="session")
def ref_modules():
    """The unmodified reference C++ (oracle/_ref), or skip when it was never built."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import aggregation_cpp  # noqa: PLC0415
        import matching_cost_cpp  # noqa: PLC0415
    except ImportError:
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    return matching_cost_cpp, aggregation_cpp
