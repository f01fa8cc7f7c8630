"""ctypes binding of the in-tree C-ABI library ``libapnerf.so`` (include/apnerf.h).

The prototypes are parsed from the header, so the header is the single source of truth for
the boundary.  There is NO CPU fallback: if the library is missing or a call fails, the
product path raises (RuntimeError), exactly like the reference's TORCH_CHECKs.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# APNERF_LIB_PATH: A/B measurements of another build of the SAME C-ABI (tools/); never a fallback
LIB_PATH = os.environ.get("APNERF_LIB_PATH") or os.path.join(_PKG_DIR, "libapnerf.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "apnerf.h")

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
}


# element types a tensor may have when it is passed for a pointer parameter of the given C type (void* = untyped
# buffers: fp16 tables / weight blobs / packed rows)
_PTR_DTYPES = {
    "float": (torch.float32,),
    "double": (torch.float64,),
    "int": (torch.int32,),
    "int64_t": (torch.int64,),
    "long long": (torch.int64,),
    "uint8_t": (torch.uint8, torch.bool),
    "uint32_t": (torch.int32, torch.uint32),
    "void": None,
    "char": None,
}
ARG_DTYPES = {}  # name -> per-argument tuple of allowed dtypes (None: not a typed pointer)


def parse_header(path: str = HEADER_PATH):
    """Return {name: (restype, [argtypes])} for every prototype declared in the header; fills ARG_DTYPES with the
    element type each pointer parameter declares, so that a tensor of another dtype raises instead of being
    reinterpreted (the reference's ``data_ptr<float>()`` raises too)."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*(const char\*|int|long long|void)\s+(apnerf_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes, dtypes = [], []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    base = a.split("*")[0].replace("const ", "").strip()
                    dtypes.append(_PTR_DTYPES.get(base))
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CTYPES[ty])
                    dtypes.append(None)
        ARG_DTYPES[name] = tuple(dtypes)
        restype = {"const char*": ctypes.c_char_p, "int": ctypes.c_int, "long long": ctypes.c_longlong,
                   "void": None}[ret]
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self._protos = None

    def _load(self):
        if self._dll is not None:
            return
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA kernels first (python __graft_entry__.py build); "
                "this package has no CPU or PyTorch fallback.")
        dll = ctypes.CDLL(LIB_PATH)
        protos = parse_header()
        for name, (restype, argtypes) in protos.items():
            fn = getattr(dll, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll, self._protos = dll, protos

    def symbols(self):
        self._load()
        return sorted(self._protos)

    def last_error(self) -> str:
        self._load()
        return (self._dll.apnerf_last_error() or b"").decode()

    def raw(self, name):
        self._load()
        return getattr(self._dll, name)

    def call(self, name, *args):
        """Call an int-returning entry point; tensors are passed as device pointers, the
        current CUDA stream is appended automatically when the prototype ends with `stream`."""
        self._load()
        fn = getattr(self._dll, name)
        conv = []
        for i, a in enumerate(args):
            if isinstance(a, torch.Tensor):
                if not a.is_cuda:
                    raise RuntimeError(f"{name}: expected a CUDA tensor, got {a.device}")
                if not a.is_contiguous():
                    raise RuntimeError(f"{name}: tensor arguments must be contiguous")
                _check_dtype(name, i, a)
                conv.append(ctypes.c_void_p(a.data_ptr()))
            elif a is None:
                conv.append(ctypes.c_void_p(0))
            else:
                conv.append(a)
        if len(conv) == len(fn.argtypes) - 1:
            conv.append(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        LAUNCHES[name] = LAUNCHES.get(name, 0) + 1
        rc = fn(*conv)
        if rc != 0:
            raise RuntimeError(f"{name} failed (code {rc}): {self.last_error()}")


class PreparedCall:
    """A C-ABI call with its arguments converted once: the per-iteration kernel sequence of the
    renderer re-launches the same entry points with the same pointers hundreds of times, and the
    ctypes argument marshalling would otherwise cost more than the small kernels themselves."""

    __slots__ = ("name", "fn", "args", "keep", "has_stream")

    def __init__(self, name, *args):
        LIB._load()
        self.name = name
        self.fn = getattr(LIB._dll, name)
        conv, keep = [], []
        for i, a in enumerate(args):
            if isinstance(a, torch.Tensor):
                if not a.is_cuda or not a.is_contiguous():
                    raise RuntimeError(f"{name}: tensor arguments must be contiguous CUDA tensors")
                _check_dtype(name, i, a)
                keep.append(a)
                conv.append(ctypes.c_void_p(a.data_ptr()))
            elif a is None:
                conv.append(ctypes.c_void_p(0))
            else:
                conv.append(a)
        self.has_stream = len(conv) == len(self.fn.argtypes) - 1
        if self.has_stream:
            conv.append(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.args, self.keep = tuple(conv), keep

    def use_current_stream(self):
        """Re-target the call at torch's current CUDA stream (the stream is marshalled once, at construction)."""
        if self.has_stream:
            self.args = self.args[:-1] + (ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),)

    def __call__(self):
        if CALL_HOOK is not None:  # measurement hook (bench.py wraps launches in CUDA events)
            return CALL_HOOK(self)
        self.invoke()

    def invoke(self):
        LAUNCHES[self.name] = LAUNCHES.get(self.name, 0) + 1
        rc = self.fn(*self.args)
        if rc != 0:
            raise RuntimeError(f"{self.name} failed (code {rc}): {LIB.last_error()}")


def _check_dtype(name, i, t):
    allowed = ARG_DTYPES.get(name, ())
    allowed = allowed[i] if i < len(allowed) else None
    if allowed is not None and t.dtype not in allowed:
        raise RuntimeError(f"{name}: argument {i} must be a tensor of dtype {' / '.join(str(d) for d in allowed)}, "
                           f"got {t.dtype} (the kernels read raw device pointers; cast at the call site)")


CALL_HOOK = None


# how many times each C-ABI entry point was invoked (bench.py turns this into kernel launches)
LAUNCHES = {}
KERNELS_PER_CALL = {"apnerf_exclusive_scan_i64": 3, "apnerf_pack_info": 5}

LIB = _Lib()
call = LIB.call


def kernel_launches() -> int:
    return sum(n * KERNELS_PER_CALL.get(name, 1) for name, n in LAUNCHES.items())


def require_cuda(*tensors):
    """The product has no CPU path: refuse anything that is not a CUDA tensor."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(f"apnerf ops need CUDA tensors (got a tensor on {t.device}); there is no CPU fallback")
