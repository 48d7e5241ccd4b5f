import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import semantic_icp_b200

    return semantic_icp_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def sicp(pkg):
    """The CUDA library; building is the driver's job (__graft_entry__.build) but do it here if it is missing."""
    if not os.path.exists(pkg.sicp.LIB_PATH):
        import subprocess

        subprocess.check_call(["make", "-C", pkg.PKG_ROOT, "-j8"])
    pkg.sicp.lib()
    return pkg.sicp
