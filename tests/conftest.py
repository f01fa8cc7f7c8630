import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def apnerf():
    import apnerf as pkg

    return pkg


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def ref_cuda():
    """The reference's own nerfacc CUDA kernels (oracle/_ref), GPU only."""
    from oracle import build_ref

    mod = build_ref.load_ref()
    if mod is None:
        pytest.skip("oracle/_ref/ref_nerfacc_cuda.so not built (needs /root/reference at build time)")
    return mod
