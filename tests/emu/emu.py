"""ctypes front end of tests/emu/recconv_emu.cu — the kernels' stage schedules run on the CPU.

TEST INFRASTRUCTURE ONLY: lets the CPU test tier check tiling / index logic of the CUDA source against the
oracle.  The product package never imports this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(ROOT, "build", "librecconv_emu.so")
MAXL = 6


class Desc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "C", "H", "W", "k", "level", "mode", "dtype", "wdtype", "has_bias")]


class Params(ctypes.Structure):
    _fields_ = [("w_down", ctypes.c_void_p), ("w_convs", ctypes.c_void_p * (MAXL + 1)), ("b_down", ctypes.c_void_p),
                ("b_convs", ctypes.c_void_p * (MAXL + 1))]


_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, "recconv_emu.cu")] + [os.path.join(ROOT, "recnext_b200", "csrc", f) for f in
                                                      ("recconv_body.cuh", "recconv_stages.cuh", "recconv_plan.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["nvcc", "-arch=sm_100a", "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                               "-shared", "-cudart", "static", "-o", SO, srcs[0]], stderr=subprocess.DEVNULL)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_recconv.restype = ctypes.c_int
    return _lib


WSO = os.path.join(ROOT, "build", "librecconv_wemu.so")
WEMU_FLAGS = []
_wlib = None


def wbuild(force=False):
    csrc = os.path.join(ROOT, "recnext_b200", "csrc")
    srcs = [os.path.join(HERE, "wemu.cu")] + [os.path.join(csrc, f) for f in
                                              ("wbody.cuh", "wstages.cuh", "wplan.h", "recconv_stages.cuh", "recconv_plan.h", "recconv_body.cuh")]
    if force or not os.path.exists(WSO) or any(os.path.getmtime(s) > os.path.getmtime(WSO) for s in srcs):
        os.makedirs(os.path.dirname(WSO), exist_ok=True)
        subprocess.check_call(["nvcc", "-arch=sm_100a", "-O1", "-std=c++17", "--expt-relaxed-constexpr", *WEMU_FLAGS, "-Xcompiler", "-fPIC",
                               "-shared", "-cudart", "static", "-o", WSO, srcs[0]], stderr=subprocess.DEVNULL)
    return WSO


def wlib():
    global _wlib
    if _wlib is None:
        _wlib = ctypes.CDLL(wbuild())
        _wlib.wemu_recconv.restype = ctypes.c_int
    return _wlib


def _np_dtype(dtype):
    return {0: np.float32, 1: np.uint16, 2: np.float16}[dtype]


def to_elem(a, dtype):
    """fp32 numpy -> storage array of the kernel element type (bf16 kept as uint16, RNE)."""
    a = np.ascontiguousarray(a, np.float32)
    if dtype == 0:
        return a
    if dtype == 2:
        return a.astype(np.float16)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return u.astype(np.uint16)


def from_elem(a, dtype):
    if dtype == 0:
        return a
    if dtype == 2:
        return a.astype(np.float32)
    return (a.astype(np.uint32) << 16).view(np.float32)


def run(x, p, mode="bilinear", gy=None, dtype=0, opts=(0, 0, 0, 0), team=False):
    """x, gy: fp32 numpy [B,C,H,W]; p: oracle RecConvParams.  Returns y (forward) or dict of grads (backward).
    team=True runs the team-resident schedule (wbody.cuh; opts = force_G, force_TW, force_NT, force_no_tma, num_sms)."""
    B, C, H, W = x.shape
    k, L = p.k, p.level
    has_bias = p.down_b is not None
    d = Desc(B, C, H, W, k, L, {"bilinear": 0, "nearest": 1}[mode], dtype, 0, int(has_bias))
    keep = []

    def ptr(a):
        a = np.ascontiguousarray(a, np.float32)
        keep.append(a)
        return a.ctypes.data

    pr = Params()
    pr.w_down = ptr(p.down_w)
    for j in range(L + 1):
        pr.w_convs[j] = ptr(p.convs_w[j])
    if has_bias:
        pr.b_down = ptr(p.down_b)
        for j in range(L + 1):
            pr.b_convs[j] = ptr(p.convs_b[j])
    xe = to_elem(x, dtype)
    out = np.zeros(x.shape, _np_dtype(dtype))
    opts = tuple(opts) + (0,) * (8 - len(opts))
    o = (ctypes.c_int * 8)(*opts)
    plan = (ctypes.c_int * 8)()
    fn = wlib().wemu_recconv if team else lib().emu_recconv
    npl = 8 if team else 7
    if gy is None:
        rc = fn(ctypes.byref(d), ctypes.byref(pr), xe.ctypes.data_as(ctypes.c_void_p), None,
                               out.ctypes.data_as(ctypes.c_void_p), None, None, 0, o, plan)
        if rc:
            raise RuntimeError(f"emu forward rc={rc}")
        return from_elem(out, dtype), list(plan)[:npl]
    ge = to_elem(gy, dtype)
    gw = np.zeros((L + 2, C, k * k), np.float32)
    gb = np.zeros((L + 2, C), np.float32)
    rc = fn(ctypes.byref(d), ctypes.byref(pr), xe.ctypes.data_as(ctypes.c_void_p), ge.ctypes.data_as(ctypes.c_void_p),
                           out.ctypes.data_as(ctypes.c_void_p), gw.ctypes.data_as(ctypes.c_void_p),
                           gb.ctypes.data_as(ctypes.c_void_p) if has_bias else None, 1, o, plan)
    if rc:
        raise RuntimeError(f"emu backward rc={rc}")
    return dict(gx=from_elem(out, dtype), down_w=gw[0].reshape(C, 1, k, k), convs_w=[gw[1 + j].reshape(C, 1, k, k) for j in range(L + 1)],
                down_b=gb[0] if has_bias else None, convs_b=[gb[1 + j] for j in range(L + 1)] if has_bias else None), list(plan)[:npl]
