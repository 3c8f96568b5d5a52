"""ctypes binding of libvmm_b200.so (the C ABI declared in include/vmm_b200.h).

There is no CPU fallback: `lib()` raises if the shared library has not been built, and
every wrapper raises `VMMError` with the library's message on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvmm_b200.so")
if os.environ.get("VMM_B200_LIB"):          # A/B builds of the same sources (tools/): never a fallback path
    LIB_PATH = os.environ["VMM_B200_LIB"]
_lib = None


class VMMError(RuntimeError):
    pass


class GemmSeg(C.Structure):
    _fields_ = [("a", C.c_void_p), ("lda", C.c_longlong), ("b", C.c_void_p), ("ldb", C.c_longlong), ("k", C.c_longlong),
                ("flags", C.c_uint)]


class PlanesC(C.Structure):
    """vmm_planes: up to 3 bf16 residual planes of one matrix."""
    _fields_ = [("p", C.c_void_p * 3), ("n", C.c_int), ("fmt", C.c_int)]


def planes_c(planes):
    """Sequence of bf16 tensors (most significant first) -> vmm_planes."""
    planes = [t for t in planes if t is not None]
    import torch
    s = PlanesC()
    s.n = len(planes)
    s.fmt = 1 if planes and planes[0].dtype == torch.float16 else 0
    for i, t in enumerate(planes):
        s.p[i] = t.data_ptr()
    return s


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VMMError(
                f"{LIB_PATH} is missing - build it with `python -m multi_view_sort_b200.build` "
                "(there is no CPU or PyTorch fallback for the VMM hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vmm_last_error.restype = C.c_char_p
        _lib.vmm_version.restype = C.c_int
    return _lib


def check(rc: int):
    if rc != 0:
        raise VMMError(f"libvmm_b200 error {rc}: {lib().vmm_last_error().decode(errors='replace')}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
