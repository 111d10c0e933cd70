"""ctypes binding of librecnext_b200.so (C ABI in include/recnext_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "librecnext_b200.so")
MAX_LEVEL = 6
ABI_VERSION = 3

F32, BF16, F16 = 0, 1, 2
BILINEAR, NEAREST = 0, 1


class RecConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "C", "H", "W", "k", "level", "mode", "dtype", "wdtype", "has_bias")]


class RecConvParams(ctypes.Structure):
    _fields_ = [
        ("w_down", ctypes.c_void_p),
        ("w_convs", ctypes.c_void_p * (MAX_LEVEL + 1)),
        ("b_down", ctypes.c_void_p),
        ("b_convs", ctypes.c_void_p * (MAX_LEVEL + 1)),
    ]


class NativeLibraryError(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    """Loads the CUDA library; raises NativeLibraryError if it has not been built (python -m recnext_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found. recnext_b200 has no CPU/PyTorch fallback; build the CUDA library with "
            "`python -m recnext_b200.build` (needs nvcc, targets sm_100a)."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.recnext_abi_version.restype = ctypes.c_int
    if L.recnext_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI mismatch: library {L.recnext_abi_version()}, binding {ABI_VERSION}; rebuild")
    L.recnext_last_error.restype = ctypes.c_char_p
    L.recconv_forward.restype = ctypes.c_int
    L.recconv_forward.argtypes = [ctypes.POINTER(RecConvDesc), ctypes.POINTER(RecConvParams), ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p]
    L.recconv_forward_workspace_bytes.restype = ctypes.c_size_t
    L.recconv_forward_workspace_bytes.argtypes = [ctypes.POINTER(RecConvDesc)]
    L.recconv_forward_ws.restype = ctypes.c_int
    L.recconv_forward_ws.argtypes = [ctypes.POINTER(RecConvDesc), ctypes.POINTER(RecConvParams), ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.recconv_backward_workspace_bytes.restype = ctypes.c_size_t
    L.recconv_backward_workspace_bytes.argtypes = [ctypes.POINTER(RecConvDesc)]
    L.recconv_backward.restype = ctypes.c_int
    L.recconv_backward.argtypes = [ctypes.POINTER(RecConvDesc), ctypes.POINTER(RecConvParams), ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                   ctypes.c_void_p]
    L.recconv_plan_describe.restype = ctypes.c_int
    L.recconv_plan_describe.argtypes = [ctypes.POINTER(RecConvDesc), ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    L.recconv_source_index.restype = ctypes.c_int
    L.recconv_source_index.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.recattn_down_forward.restype = ctypes.c_int
    L.recattn_down_forward.argtypes = [ctypes.POINTER(RecConvDesc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
    L.recattn_up_forward.restype = ctypes.c_int
    L.recattn_up_forward.argtypes = [ctypes.POINTER(RecConvDesc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    L.recnext_linattn_forward.restype = ctypes.c_int
    L.recnext_linattn_forward.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 5
    L.recnext_linattn_forward_qk.restype = ctypes.c_int
    L.recnext_linattn_forward_qk.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 8
    L.recnext_linattn_forward_pe.restype = ctypes.c_int
    L.recnext_linattn_forward_pe.argtypes = [ctypes.c_int32] * 6 + [ctypes.c_void_p] * 9
    L.recnext_stem_forward.restype = ctypes.c_int
    L.recnext_stem_forward.argtypes = [ctypes.c_int32] * 6 + [ctypes.c_void_p] * 7
    L.recnext_dwdown_forward.restype = ctypes.c_int
    L.recnext_dwdown_forward.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 5
    L.recnext_ffn_forward.restype = ctypes.c_int
    L.recnext_ffn_forward.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 8
    L.recnext_ffn_packed_bytes.restype = ctypes.c_size_t
    L.recnext_ffn_packed_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.recnext_ffn_pack.restype = ctypes.c_int
    L.recnext_ffn_pack.argtypes = [ctypes.c_int32] * 3 + [ctypes.c_void_p] * 4
    L.recnext_ffn_forward_packed.restype = ctypes.c_int
    L.recnext_ffn_forward_packed.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_void_p] * 7
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().recnext_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


EXPORTS = [
    "recnext_abi_version", "recnext_last_error", "recconv_forward", "recconv_forward_workspace_bytes", "recconv_forward_ws",
    "recconv_backward_workspace_bytes", "recconv_backward",
    "recconv_plan_describe", "recconv_source_index", "recattn_down_forward", "recattn_up_forward", "recnext_ffn_forward", "recnext_ffn_packed_bytes", "recnext_ffn_pack", "recnext_ffn_forward_packed", "recnext_dwdown_forward", "recnext_stem_forward", "recnext_linattn_forward", "recnext_linattn_forward_qk", "recnext_linattn_forward_pe",
]
