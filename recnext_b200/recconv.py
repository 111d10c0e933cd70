"""RecConv2d — drop-in for the reference module (reference model/recnext.py:8-34), backed by fused sm_100a CUDA.

Same constructor, same sub-modules and therefore the same ``state_dict`` keys and shapes
(``down.weight``, ``convs.{j}.weight`` [C,1,k,k], optional ``.bias`` [C]); ``forward`` runs ONE fused kernel that
keeps the whole pyramid in shared memory (x is read once, y written once) and ``backward`` runs one fused
kernel plus a tiny deterministic reduction of the filter gradients.  CUDA tensors only: anything else raises
(no CPU fallback, no cuDNN dispatch).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.nn as nn

from . import _native as N

_DTYPES = {torch.float32: N.F32, torch.bfloat16: N.BF16, torch.float16: N.F16}
_MODES = {"bilinear": N.BILINEAR, "nearest": N.NEAREST}


def _desc(x: torch.Tensor, k: int, level: int, mode: str, wdtype: torch.dtype, has_bias: bool) -> N.RecConvDesc:
    if x.dim() != 4:
        raise ValueError(f"RecConv2d expects [B,C,H,W], got {tuple(x.shape)}")
    if x.dtype not in _DTYPES:
        raise TypeError(f"RecConv2d: unsupported dtype {x.dtype} (float32, bfloat16, float16)")
    if mode not in _MODES:
        raise ValueError(f"RecConv2d: unsupported interpolation mode {mode!r} (bilinear, nearest)")
    B, C, H, W = x.shape
    return N.RecConvDesc(B, C, H, W, k, level, _MODES[mode], _DTYPES[x.dtype], _DTYPES[wdtype], int(has_bias))


def _params(weights: List[torch.Tensor], biases: Optional[List[torch.Tensor]]) -> N.RecConvParams:
    """weights/biases: [down, convs[0], ..., convs[L]] (contiguous CUDA tensors)."""
    p = N.RecConvParams()
    p.w_down = weights[0].data_ptr()
    for j, w in enumerate(weights[1:]):
        p.w_convs[j] = w.data_ptr()
    if biases is not None:
        p.b_down = biases[0].data_ptr()
        for j, b in enumerate(biases[1:]):
            p.b_convs[j] = b.data_ptr()
    return p


def _stream(x: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def _prep_params(x, weights, biases):
    wd = weights[0].dtype
    if wd != torch.float32 and wd != x.dtype:
        raise TypeError(f"RecConv2d: parameters must be float32 or match the input dtype ({wd} vs {x.dtype})")
    ws = [w.detach().contiguous() for w in weights]
    bs = [b.detach().contiguous() for b in biases] if biases is not None else None
    for t in ws + (bs or []):
        if t.device != x.device or t.dtype != wd:
            raise ValueError("RecConv2d: all parameters must live on the input's device and share one dtype")
    return wd, ws, bs


# ---- optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream -------------
_timing = None


def timing_begin() -> None:
    """Start bracketing every forward launch with CUDA events recorded on the launching stream."""
    global _timing
    _timing = []


def timing_end():
    """Stop; returns [{'bytes': algorithmic bytes (2*N*e), 'ms': device time, 'shape': ...}] per launch."""
    global _timing
    rec, _timing = _timing or [], None
    torch.cuda.synchronize()
    return [dict(bytes=b, ms=a.elapsed_time(z), shape=shape) for (a, z, b, shape) in rec]


def recconv_forward(x: torch.Tensor, weights: List[torch.Tensor], biases: Optional[List[torch.Tensor]], k: int, level: int,
                    mode: str) -> torch.Tensor:
    """Functional forward.  weights = [down.weight, convs[0].weight, ..., convs[level].weight]."""
    if not x.is_cuda:
        raise RuntimeError("recnext_b200.RecConv2d runs on CUDA (sm_100a) only; there is no CPU fallback")
    x = x.contiguous()
    wd, ws, bs = _prep_params(x, weights, biases)
    d = _desc(x, k, level, mode, wd, bs is not None)
    y = torch.empty_like(x)
    # 0 for every plane whose pyramid fits on chip (the fused kernels); planes that do not fit are streamed through a workspace
    nbytes = N.lib().recconv_forward_workspace_bytes(ctypes.byref(d))
    ws_buf = torch.empty((nbytes,), dtype=torch.uint8, device=x.device) if nbytes else None
    with torch.cuda.device(x.device):
        if _timing is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        N.check(N.lib().recconv_forward_ws(ctypes.byref(d), ctypes.byref(_params(ws, bs)), x.data_ptr(), y.data_ptr(),
                                           ws_buf.data_ptr() if nbytes else None, nbytes, _stream(x)), "recconv_forward")
        if _timing is not None:
            ev1.record()
            _timing.append((ev0, ev1, 2 * x.numel() * x.element_size(), tuple(x.shape)))
    return y


def recconv_backward(x, gy, weights, biases, k, level, mode):
    """-> (gx, gw [(L+2),C,k*k] fp32, gb [(L+2),C] fp32 or None); slot 0 = down, 1+j = convs[j]."""
    if not x.is_cuda:
        raise RuntimeError("recnext_b200.RecConv2d runs on CUDA (sm_100a) only; there is no CPU fallback")
    x = x.contiguous()
    gy = gy.contiguous()
    if gy.dtype != x.dtype:
        gy = gy.to(x.dtype)
    wd, ws, bs = _prep_params(x, weights, biases)
    d = _desc(x, k, level, mode, wd, bs is not None)
    C = x.shape[1]
    gx = torch.empty_like(x)
    gw = torch.empty((level + 2, C, k * k), dtype=torch.float32, device=x.device)
    gb = torch.empty((level + 2, C), dtype=torch.float32, device=x.device) if bs is not None else None
    nbytes = N.lib().recconv_backward_workspace_bytes(ctypes.byref(d))
    if nbytes == 0 and x.numel() > 0:
        N.check(N.lib().recconv_plan_describe(ctypes.byref(d), 1, ctypes.create_string_buffer(8), 8), "recconv_backward(plan)")
    ws_buf = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().recconv_backward(ctypes.byref(d), ctypes.byref(_params(ws, bs)), x.data_ptr(), gy.data_ptr(), gx.data_ptr(),
                                         gw.data_ptr(), gb.data_ptr() if gb is not None else None, ws_buf.data_ptr(), nbytes,
                                         _stream(x)), "recconv_backward")
    return gx, gw, gb


def plan_describe(shape, k=5, level=2, mode="bilinear", dtype=torch.float32, bias=False, backward=False) -> str:
    B, C, H, W = shape
    d = N.RecConvDesc(B, C, H, W, k, level, _MODES[mode], _DTYPES[dtype], N.F32, int(bias))
    buf = ctypes.create_string_buffer(512)
    N.check(N.lib().recconv_plan_describe(ctypes.byref(d), int(backward), buf, 512), "recconv_plan_describe")
    return buf.value.decode()


class _RecConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, level, mode, nb, *params):
        weights = list(params[: level + 2])
        biases = list(params[level + 2:]) if nb else None
        ctx.save_for_backward(x, *params)
        ctx.cfg = (k, level, mode, nb)
        return recconv_forward(x, weights, biases, k, level, mode)

    @staticmethod
    def backward(ctx, gy):
        x, *params = ctx.saved_tensors
        k, level, mode, nb = ctx.cfg
        weights = list(params[: level + 2])
        biases = list(params[level + 2:]) if nb else None
        gx, gw, gb = recconv_backward(x, gy, weights, biases, k, level, mode)
        C = x.shape[1]
        wdt = weights[0].dtype
        gws = [gw[j].view(C, 1, k, k).to(wdt) for j in range(level + 2)]
        if level == 0:
            gws[0] = None  # `down` does not take part in the graph at level 0 (reference: grad stays None)
        grads = [gx if ctx.needs_input_grad[0] else None, None, None, None, None] + gws
        if nb:
            gbs = [gb[j].to(wdt) for j in range(level + 2)]
            if level == 0:
                gbs[0] = None
            grads += gbs
        return tuple(grads)


class RecConv2d(nn.Module):
    """``RecConv2d(in_channels, kernel_size=5, bias=False, level=2, mode='bilinear')`` — reference model/recnext.py:9."""

    def __init__(self, in_channels, kernel_size=5, bias=False, level=2, mode="bilinear"):
        super().__init__()
        if kernel_size not in (3, 5, 7):
            raise ValueError(f"RecConv2d: kernel_size {kernel_size} not supported by the CUDA kernels (3, 5, 7)")
        if not 0 <= level <= N.MAX_LEVEL:
            raise ValueError(f"RecConv2d: level {level} outside 0..{N.MAX_LEVEL}")
        if mode not in _MODES:
            raise ValueError(f"RecConv2d: mode {mode!r} not supported (bilinear, nearest)")
        self.level = level
        self.mode = mode
        self.kernel_size = kernel_size
        kwargs = dict(in_channels=in_channels, out_channels=in_channels, groups=in_channels, kernel_size=kernel_size,
                      padding=kernel_size // 2, bias=bias)
        # parameter containers only (never called): keeps state_dict / init / .to() / DDP identical to the reference
        self.down = nn.Conv2d(stride=2, **kwargs)
        self.convs = nn.ModuleList([nn.Conv2d(**kwargs) for _ in range(level + 1)])

    def _param_lists(self):
        mods = [self.down, *self.convs]
        weights = [m.weight for m in mods]
        biases = [m.bias for m in mods] if self.down.bias is not None else []
        return weights, biases

    def forward(self, x):
        weights, biases = self._param_lists()
        if torch.is_autocast_enabled() and x.is_cuda:
            x = x.to(torch.get_autocast_dtype("cuda"))  # conv2d is an autocast-to-low-precision op in the reference path
        # the kernels work on NCHW planes; a channels_last input is transposed once and the result is handed back in
        # the caller's memory format (what nn.Conv2d does), so a channels_last model keeps its 1x1 convs transpose-free
        channels_last = x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)
        if x.is_cuda:
            # the registered operator (recnext_b200/ops.py): one node for jit.trace / torch.compile, same C-ABI call in eager
            from . import ops as _ops  # noqa: F401  (registers torch.ops.recnext.*)

            y = torch.ops.recnext.recconv(x, list(weights), list(biases), self.kernel_size, self.level, self.mode)
        else:
            y = _RecConvFn.apply(x, self.kernel_size, self.level, self.mode, len(biases), *weights, *biases)   # raises: no CPU path
        return y.contiguous(memory_format=torch.channels_last) if channels_last else y

    def extra_repr(self):
        return f"level={self.level}, mode={self.mode!r}, kernel_size={self.kernel_size}"
