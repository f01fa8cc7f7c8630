"""TEST INFRASTRUCTURE ONLY (oracle).

Builds the reference's own nerfacc CUDA kernels (grid.cu, scan.cu) -- compiled IN PLACE from
/root/reference, never copied -- into oracle/_ref/ref_nerfacc_cuda.so.  The result is the
bit-exact checker for the ray-march kernels and the tolerance checker for the packed scan
(SURVEY.md section 8c).  It needs a GPU to run, so it is only used by `-m gpu` tests.

Recipe (what a Makefile would do): nvcc -O3 (the reference's own flags,
perception/nerfacc/nerfacc/cuda/_backend.py:43-44) for sm_100a on
  <ref>/perception/nerfacc/nerfacc/cuda/csrc/grid.cu
  <ref>/perception/nerfacc/nerfacc/cuda/csrc/scan.cu
plus oracle/ref_binding.cpp (our pybind stub for the four hot-path symbols).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/perception/nerfacc/nerfacc/cuda/csrc"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "ref_nerfacc_cuda"


def ref_so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    if os.path.exists(ref_so_path()):
        return ref_so_path()
    if not os.path.isdir(REF_CSRC):
        return None  # GPU box: only the prebuilt file is used
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    build_dir = os.path.join(OUT_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    load(
        name=NAME,
        sources=[
            os.path.join(REF_CSRC, "grid.cu"),
            os.path.join(REF_CSRC, "scan.cu"),
            os.path.join(HERE, "ref_binding.cpp"),
        ],
        extra_include_paths=[REF_CSRC],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3"],
        build_directory=build_dir,
        is_python_module=False,
        verbose=verbose,
    )
    shutil.copy(os.path.join(build_dir, NAME + ".so"), ref_so_path())
    shutil.rmtree(build_dir, ignore_errors=True)
    return ref_so_path()


def load_ref():
    """Import the prebuilt reference module (GPU tests only)."""
    import importlib.util

    import torch  # noqa: F401  (must be loaded before the extension)

    path = ref_so_path()
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose=True)
    print("reference oracle:", p)
    sys.exit(0 if p else 1)
