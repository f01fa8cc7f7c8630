import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The suite needs libapnerf.so (the product refuses to run without it).  It normally exists -- build() made it,
    and it travels to the GPU box -- but a fresh checkout has only sources: compile it once here (nvcc cross-compiles
    sm_100a without a GPU)."""
    pkg = os.path.join(ROOT, "active-perception-using-neural-radiance-fields_b200")
    so = os.path.join(pkg, "libapnerf.so")
    csrc = os.path.join(pkg, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs)
    if stale and os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):  # incremental: seconds when little changed
        import subprocess

        subprocess.run([sys.executable, os.path.join(csrc, "build.py")], check=True, cwd=csrc)
    yield


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
