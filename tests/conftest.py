import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The reference's own C++ compiled headless (oracle/_ref/libbpt_ref.so).  Test infrastructure only."""
    from oracle import ref_oracle
    if not ref_oracle.available():
        ref_oracle.build()
    if not ref_oracle.available():
        pytest.skip("oracle/_ref/libbpt_ref.so not built and /root/reference absent")
    return ref_oracle


@pytest.fixture(scope="session")
def bpt():
    import buas_pathtracer_b200 as B
    if not os.path.exists(B.library_path()):
        B.build_library()
    B.load_library()
    return B


@pytest.fixture(scope="session")
def renderer(bpt):
    return bpt.Renderer(0)
