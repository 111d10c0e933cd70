"""ctypes/numpy front end of oracle/recconv_oracle.c.

TEST INFRASTRUCTURE ONLY (see the C file's header): the checker for tests/, smoke() and
bench.py's cpu_baseline leg.  The product package never imports this module.

`RecConvParams` mirrors the reference module's state_dict layout
(/root/reference/model/recnext.py:9-22): down.weight [C,1,k,k], convs.{j}.weight [C,1,k,k],
optional down.bias / convs.{j}.bias [C].
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librecconv_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (a couple of seconds)."""
    src = os.path.join(_HERE, "recconv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.recconv_oracle_forward.restype = ctypes.c_int
        _lib.recconv_oracle_backward.restype = ctypes.c_int
        _lib.recconv_oracle_nearest_index.restype = ctypes.c_int
        _lib.recconv_oracle_pyramid.restype = ctypes.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


@dataclass
class RecConvParams:
    """Weights of one RecConv2d, numpy fp32, in reference state_dict shapes."""

    down_w: np.ndarray  # [C,1,k,k]
    convs_w: List[np.ndarray]  # (L+1) × [C,1,k,k]
    down_b: Optional[np.ndarray] = None  # [C]
    convs_b: Optional[List[np.ndarray]] = None  # (L+1) × [C]

    @property
    def level(self) -> int:
        return len(self.convs_w) - 1

    @property
    def k(self) -> int:
        return int(self.down_w.shape[-1])

    @classmethod
    def from_state_dict(cls, sd, prefix: str = "") -> "RecConvParams":
        g = lambda n: np.ascontiguousarray(sd[prefix + n].detach().float().cpu().numpy())  # noqa: E731
        L = 0
        while f"{prefix}convs.{L + 1}.weight" in sd:
            L += 1
        has_b = f"{prefix}down.bias" in sd
        return cls(
            down_w=g("down.weight"),
            convs_w=[g(f"convs.{j}.weight") for j in range(L + 1)],
            down_b=g("down.bias") if has_b else None,
            convs_b=[g(f"convs.{j}.bias") for j in range(L + 1)] if has_b else None,
        )

    @classmethod
    def random(cls, C: int, k: int, level: int, bias: bool, rng: np.random.Generator, bound: Optional[float] = None):
        """nn.Conv2d default init for depthwise: U(-1/sqrt(k*k), 1/sqrt(k*k)) for weight and bias."""
        b = bound if bound is not None else 1.0 / k
        u = lambda *s: rng.uniform(-b, b, size=s).astype(np.float32)  # noqa: E731
        return cls(
            down_w=u(C, 1, k, k),
            convs_w=[u(C, 1, k, k) for _ in range(level + 1)],
            down_b=u(C) if bias else None,
            convs_b=[u(C) for _ in range(level + 1)] if bias else None,
        )

    def packed(self):
        wc = np.ascontiguousarray(np.stack([w.reshape(w.shape[0], -1) for w in self.convs_w]).astype(np.float32))
        wd = np.ascontiguousarray(self.down_w.reshape(self.down_w.shape[0], -1).astype(np.float32))
        bd = None if self.down_b is None else np.ascontiguousarray(self.down_b.astype(np.float32))
        bc = None if self.convs_b is None else np.ascontiguousarray(np.stack(self.convs_b).astype(np.float32))
        return wd, bd, wc, bc


MODES = {"bilinear": 0, "nearest": 1}


def forward(x: np.ndarray, p: RecConvParams, mode: str = "bilinear", round_bf16: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, C, H, W = x.shape
    wd, bd, wc, bc = p.packed()
    y = np.empty_like(x)
    rc = lib().recconv_oracle_forward(_p(x), _p(wd), _p(bd), _p(wc), _p(bc), _p(y), B, C, H, W, p.k, p.level,
                                      MODES[mode], int(round_bf16))
    if rc:
        raise ValueError(f"recconv_oracle_forward: bad argument (code {rc})")
    return y


def backward(x: np.ndarray, gy: np.ndarray, p: RecConvParams, mode: str = "bilinear"):
    """Returns dict(gx, down_w, down_b, convs_w[list], convs_b[list]) — grads in state_dict shapes."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    gy = np.ascontiguousarray(gy, dtype=np.float32)
    B, C, H, W = x.shape
    k, L = p.k, p.level
    wd, bd, wc, bc = p.packed()
    gx = np.empty_like(x)
    gwd = np.empty((C, k * k), np.float32)
    gwc = np.empty((L + 1, C, k * k), np.float32)
    gbd = np.empty((C,), np.float32) if bd is not None else None
    gbc = np.empty((L + 1, C), np.float32) if bc is not None else None
    rc = lib().recconv_oracle_backward(_p(x), _p(gy), _p(wd), _p(bd), _p(wc), _p(bc), _p(gx), _p(gwd), _p(gbd),
                                       _p(gwc), _p(gbc), B, C, H, W, k, L, MODES[mode])
    if rc:
        raise ValueError(f"recconv_oracle_backward: bad argument (code {rc})")
    return dict(
        gx=gx,
        down_w=gwd.reshape(C, 1, k, k),
        down_b=gbd,
        convs_w=[gwc[j].reshape(C, 1, k, k) for j in range(L + 1)],
        convs_b=None if gbc is None else [gbc[j] for j in range(L + 1)],
    )


def bilinear_index(in_size: int, out_size: int, dst: int):
    i0, i1, lam = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
    lib().recconv_oracle_bilinear_index(in_size, out_size, dst, ctypes.byref(i0), ctypes.byref(i1), ctypes.byref(lam))
    return i0.value, i1.value, lam.value


def nearest_index(in_size: int, out_size: int, dst: int) -> int:
    return lib().recconv_oracle_nearest_index(in_size, out_size, dst)


def pyramid(H: int, W: int, k: int, level: int):
    Hs = (ctypes.c_int * (level + 1))()
    Ws = (ctypes.c_int * (level + 1))()
    rc = lib().recconv_oracle_pyramid(H, W, k, level, Hs, Ws)
    if rc:
        raise ValueError("bad level")
    return list(Hs), list(Ws)
