import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference headers on the Eigen shim, if buildable/present."""
    from oracle.binding import load_ref
    r = load_ref()
    if r is None:
        pytest.skip("oracle/_ref not available (no /root/reference and no prebuilt copy)")
    return r


@pytest.fixture(scope="session")
def wb():
    import walnuts_b200
    return walnuts_b200
