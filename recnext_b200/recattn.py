"""RecAttn2d — drop-in for the reference A-series token mixer (reference model/recattn.py:54-67).

    forward(x) = conv(x + interpolate(LinearAttention(down_conv(x)), size=x.shape[2:], mode))

Same constructor, sub-module names and ``state_dict`` layout as the reference (``down.0`` ConvNorm k x k stride 2,
``down.1`` LinearAttention1/2 with ``qk`` / ``pe`` ConvNorms, ``conv`` ConvNorm), including ``ConvNorm.fuse()``
(model/recattn.py:87-111) so fused-BN eval checkpoints load unchanged.

What runs where:
  * the two plane-independent, memory-bound pieces — the stride-2 depthwise conv and the fused
    ``conv(x + interpolate(z))`` tail — run in the hand-written sm_100a tensor-core kernels behind
    ``recattn_down_forward`` / ``recattn_up_forward`` (include/recnext_b200.h), BatchNorm folded into (w, b);
  * the linear attention in between: its contraction chain (elu + 1, k v^T, mean(k), q^T kv / (q^T mean(k) + 1e-6), + pe;
    model/recattn.py:21-28, 44-51) is one more kernel (``recnext_linattn_forward``); its two ConvNorms (grouped 1x1 ``qk``,
    depthwise 3x3 ``pe``) stay library convs.
Inference only (CUDA activations, eval mode; the two RecAttn2d pieces 16-bit): training through the custom kernels needs
their backward, which is not built, so ``forward`` raises in training mode, on CPU tensors and for head sizes the kernel does
not have — there is no PyTorch fallback anywhere in this file (the eager restatement lives in oracle/torch_ref.py, for tests).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _native as N
from . import recconv as _rc
from .recconv import _DTYPES, _MODES, _stream


def _call_desc(x: torch.Tensor, mode: str, wdtype: torch.dtype, has_bias: bool) -> N.RecConvDesc:
    if x.dim() != 4:
        raise ValueError(f"RecAttn2d expects [B,C,H,W], got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("recnext_b200.RecAttn2d runs on CUDA (sm_100a) only; there is no CPU fallback")
    if x.dtype not in _DTYPES:
        raise TypeError(f"RecAttn2d: float32 / bfloat16 / float16 activations only, got {x.dtype}")   # 16-bit: tensor-core kernels; fp32: the 1e-5 path
    B, C, H, W = x.shape
    return N.RecConvDesc(B, C, H, W, 5, 1, _MODES[mode], _DTYPES[x.dtype], _DTYPES[wdtype], int(has_bias))


def _timing_start():
    """bench.py's per-launch timing (recconv.timing_begin/end): CUDA events on the launching stream"""
    if _rc._timing is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _timing_stop(ev0, alg_bytes, shape):
    if ev0 is None or _rc._timing is None:
        return
    ev1 = torch.cuda.Event(enable_timing=True)
    ev1.record()
    _rc._timing.append((ev0, ev1, alg_bytes, shape))


def _prep(x, w, b):
    if w.dtype != torch.float32 and w.dtype != x.dtype:
        raise TypeError(f"RecAttn2d: parameters must be float32 or match the input dtype ({w.dtype} vs {x.dtype})")
    w = w.detach().contiguous()
    b = None if b is None else b.detach().to(w.dtype).contiguous()
    if tuple(w.shape) != (x.shape[1], 1, 5, 5):
        raise ValueError(f"RecAttn2d: depthwise 5x5 filter expected, got {tuple(w.shape)}")
    return w, b


def recattn_down_forward(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """depthwise 5x5 stride-2 conv (+ bias) — RecAttn2d.down[0] with its BatchNorm folded (model/recattn.py:60)."""
    x = x.contiguous()
    d = _call_desc(x, "nearest", w.dtype, b is not None)
    w, b = _prep(x, w, b)
    B, C, H, W = x.shape
    out = torch.empty(B, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1, device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        ev = _timing_start()
        N.check(N.lib().recattn_down_forward(ctypes.byref(d), w.data_ptr(), None if b is None else b.data_ptr(), x.data_ptr(),
                                             out.data_ptr(), _stream(x)), "recattn_down_forward")
        _timing_stop(ev, (x.numel() + out.numel()) * x.element_size(), ("down",) + tuple(x.shape))
    return out


def recattn_up_forward(x: torch.Tensor, z: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], mode: str = "nearest") -> torch.Tensor:
    """conv(x + interpolate(z, size=x.shape[2:], mode)) — the tail of RecAttn2d.forward (model/recattn.py:67)."""
    x = x.contiguous()
    z = z.to(x.dtype).contiguous()
    if z.dim() != 4 or z.shape[:2] != x.shape[:2]:
        raise ValueError(f"RecAttn2d: z {tuple(z.shape)} does not match x {tuple(x.shape)}")
    d = _call_desc(x, mode, w.dtype, b is not None)
    w, b = _prep(x, w, b)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ev = _timing_start()
        N.check(N.lib().recattn_up_forward(ctypes.byref(d), w.data_ptr(), None if b is None else b.data_ptr(), x.data_ptr(), z.data_ptr(),
                                           int(z.shape[2]), int(z.shape[3]), y.data_ptr(), _stream(x)), "recattn_up_forward")
        _timing_stop(ev, (2 * x.numel() + z.numel()) * x.element_size(), ("up",) + tuple(x.shape))
    return y


class ConvNorm(nn.Sequential):
    """Conv2d + BatchNorm2d with the reference's fuse() (model/recattn.py:70-111)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, groups=1, bias=False, bn_weight_init=1):
        super().__init__()
        self.add_module("conv", nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias=bias))
        self.add_module("norm", nn.BatchNorm2d(out_channels))
        nn.init.constant_(self.norm.weight, bn_weight_init)
        nn.init.constant_(self.norm.bias, 0)

    @torch.no_grad()
    def folded(self):
        """(w, b) of the equivalent biased conv (eval-mode BatchNorm folded), fp32"""
        s = self.norm.weight / (self.norm.running_var + self.norm.eps) ** 0.5
        b = self.norm.bias - s * self.norm.running_mean
        if self.conv.bias is not None:
            b = b + s * self.conv.bias
        return (s[:, None, None, None] * self.conv.weight).float(), b.float()

    @torch.no_grad()
    def fuse(self):
        w, b = self.folded()
        c = self.conv
        m = nn.Conv2d(w.size(1) * c.groups, w.size(0), w.shape[2:], stride=c.stride, padding=c.padding, dilation=c.dilation,
                      groups=c.groups, device=c.weight.device)
        m.weight.data.copy_(w)
        m.bias.data.copy_(b)
        return m


def _wb(m):
    """(w, b) of a ConvNorm (BatchNorm folded on the fly) or of the nn.Conv2d that fuse() left in its place"""
    if isinstance(m, ConvNorm):
        return m.folded()
    return m.weight, m.bias


def linattn_forward(qk: torch.Tensor, v: torch.Tensor, pe: Optional[torch.Tensor], num_heads: int) -> torch.Tensor:
    """q, k = elu(qk) + 1;  out = q^T (k v^T / n) / (q^T mean(k) + 1e-6) (+ pe) per image and head — everything between the
    ``qk`` ConvNorm and the final add of LinearAttention1/2.forward (reference model/recattn.py:21-28, :44-51) as ONE sm_100a
    kernel (``recnext_linattn_forward``).  qk: [B, 2*dim, h, w] pre-activation, v / pe: [B, dim, h, w]; CUDA tensors, fp32 or 16-bit."""
    if not qk.is_cuda:
        raise RuntimeError("recnext_b200.linattn_forward runs on CUDA (sm_100a) only; there is no CPU fallback")
    if qk.dtype not in _DTYPES:
        raise TypeError(f"linattn_forward: float32 / bfloat16 / float16 only, got {qk.dtype}")
    qk, v = qk.contiguous(), v.to(qk.dtype).contiguous()
    pe = None if pe is None else pe.to(qk.dtype).contiguous()
    B, dim, H, W = v.shape
    if qk.shape[1] != 2 * dim or tuple(qk.shape[2:]) != (H, W):
        raise ValueError(f"linattn_forward: qk {tuple(qk.shape)} does not match v {tuple(v.shape)}")
    out = torch.empty_like(v)
    with torch.cuda.device(v.device):
        ev = _timing_start()
        N.check(N.lib().recnext_linattn_forward(B, dim, num_heads, H * W, _DTYPES[qk.dtype], qk.data_ptr(), v.data_ptr(),
                                                None if pe is None else pe.data_ptr(), out.data_ptr(), _stream(v)), "recnext_linattn_forward")
        _timing_stop(ev, (qk.numel() + 2 * v.numel() + (0 if pe is None else pe.numel())) * v.element_size(), ("linattn",) + tuple(v.shape))
    return out


def linattn_forward_qk(q: torch.Tensor, k: torch.Tensor, qbias: Optional[torch.Tensor], kbias: Optional[torch.Tensor], v: torch.Tensor,
                       pe: Optional[torch.Tensor], num_heads: int) -> torch.Tensor:
    """``linattn_forward`` with q and k as two [B, dim, h, w] (or [B, dim, n]) pre-activation tensors and optional fp32 biases [dim] that the
    kernel adds before the elu (``recnext_linattn_forward_qk``): the caller's GEMMs need no bias pass."""
    if not v.is_cuda:
        raise RuntimeError("recnext_b200.linattn_forward_qk runs on CUDA (sm_100a) only; there is no CPU fallback")
    if v.dtype not in _DTYPES:
        raise TypeError(f"linattn_forward_qk: float32 / bfloat16 / float16 only, got {v.dtype}")
    v = v.contiguous()
    q, k = q.to(v.dtype).contiguous(), k.to(v.dtype).contiguous()
    pe = None if pe is None else pe.to(v.dtype).contiguous()
    B, dim, H, W = v.shape
    if q.numel() != v.numel() or k.numel() != v.numel():
        raise ValueError(f"linattn_forward_qk: q {tuple(q.shape)} / k {tuple(k.shape)} do not match v {tuple(v.shape)}")
    qb = None if qbias is None else qbias.detach().float().contiguous()
    kb = None if kbias is None else kbias.detach().float().contiguous()
    out = torch.empty_like(v)
    with torch.cuda.device(v.device):
        ev = _timing_start()
        N.check(N.lib().recnext_linattn_forward_qk(B, dim, num_heads, H * W, _DTYPES[v.dtype], q.data_ptr(), k.data_ptr(),
                                                   None if qb is None else qb.data_ptr(), None if kb is None else kb.data_ptr(), v.data_ptr(),
                                                   None if pe is None else pe.data_ptr(), out.data_ptr(), _stream(v)), "recnext_linattn_forward_qk")
        _timing_stop(ev, (4 * v.numel() + (0 if pe is None else pe.numel())) * v.element_size(), ("linattn",) + tuple(v.shape))
    return out


def linattn_forward_pe(q: torch.Tensor, k: torch.Tensor, qbias, kbias, v: torch.Tensor, pe_w: torch.Tensor, pe_b: Optional[torch.Tensor],
                       num_heads: int) -> torch.Tensor:
    """``linattn_forward_qk`` with the depthwise 3x3 ``pe`` ConvNorm (BatchNorm folded: pe_w [dim, 1, 3, 3], pe_b [dim]) evaluated on v inside the
    kernel (``recnext_linattn_forward_pe``): everything of LinearAttention1/2.forward after the ``qk`` GEMMs is one launch."""
    if not v.is_cuda:
        raise RuntimeError("recnext_b200.linattn_forward_pe runs on CUDA (sm_100a) only; there is no CPU fallback")
    if v.dtype not in _DTYPES:
        raise TypeError(f"linattn_forward_pe: float32 / bfloat16 / float16 only, got {v.dtype}")
    v = v.contiguous()
    q, k = q.to(v.dtype).contiguous(), k.to(v.dtype).contiguous()
    B, dim, H, W = v.shape
    if q.numel() != v.numel() or k.numel() != v.numel() or pe_w.numel() != dim * 9:
        raise ValueError(f"linattn_forward_pe: q {tuple(q.shape)} / k {tuple(k.shape)} / pe_w {tuple(pe_w.shape)} do not match v {tuple(v.shape)}")
    f32 = lambda t: None if t is None else t.detach().float().contiguous()  # noqa: E731
    qb, kb, pw, pb = f32(qbias), f32(kbias), f32(pe_w), f32(pe_b)
    ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    out = torch.empty_like(v)
    with torch.cuda.device(v.device):
        ev = _timing_start()
        N.check(N.lib().recnext_linattn_forward_pe(B, dim, num_heads, H, W, _DTYPES[v.dtype], q.data_ptr(), k.data_ptr(), ptr(qb), ptr(kb), v.data_ptr(),
                                                   ptr(pw), ptr(pb), out.data_ptr(), _stream(v)), "recnext_linattn_forward_pe")
        _timing_stop(ev, 4 * v.numel() * v.element_size(), ("linattn",) + tuple(v.shape))
    return out


def _no_grad_wanted(x, module, what):
    """The A-series kernels have no backward: refuse to cut a gradient silently (an eval-mode model stays differentiable in the reference)."""
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        raise RuntimeError(f"recnext_b200 {what}: the sm_100a kernels have no backward and a gradient can be asked for here (grad mode is on and the "
                           "input or a parameter requires grad) — run under torch.no_grad() / inference_mode(), or freeze the parameters")


LINATTN_HEAD_DIMS = (4, 8, 16, 20, 24, 28, 32, 40)   # 20..40: the RecNeXt-A models; 4, 8, 16: small test models
# RECNEXT_LINATTN_GEMM=0: the grouped 1x1 `qk` ConvNorm runs as the library's grouped conv (round-2 path; A/B measurements)
LINATTN_GEMM = os.environ.get("RECNEXT_LINATTN_GEMM", "1") != "0"


class _LinearAttention(nn.Module):
    """The linear attention of the A-series token mixer.  The reference spells it twice — LinearAttention1 (model/recattn.py:8-28, the
    d x d ``kv`` form, stages 0-2) and LinearAttention2 (:31-51, the n x n form, stage 3, n = 16) — and checks that the two agree
    (lsnet/model/recattn.py:480-501); both are served by ONE kernel that always takes the d x d route (``recnext_linattn_forward``):
    q, k = elu(qk(x)) + 1;  out = q^T (k v^T / n) / (q^T mean(k) + 1e-6) + pe(x).  Sub-module names (``qk``, ``pe``) and therefore the
    ``state_dict`` are the reference's.  The grouped (groups = 2) 1x1 ``qk`` ConvNorm is two plain GEMMs per image, W_q x[: dim/2] and
    W_k x[dim/2 :]: the library's grouped-conv kernels for it are SGEMM-class (15 % of an A3 step with their bias adds), so it runs as two
    batched tensor-core GEMMs (``torch.bmm`` on the BatchNorm-folded weights, cuBLAS) whose bias the attention kernel adds before the elu; the
    depthwise 3x3 ``pe`` ConvNorm is evaluated on v inside the attention kernel (``recnext_linattn_forward_pe``).
    Inference only: there is no backward and, deliberately, no PyTorch fallback — anything the kernel does not serve raises."""

    def __init__(self, dim, num_heads, conv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.qk = ConvNorm(dim, dim * 2, kernel_size=1, groups=2, bias=conv_bias)   # (conv_bias: the L-series ConvNorm keeps the conv's bias,
        self.pe = ConvNorm(dim, dim, kernel_size=3, padding=1, groups=dim, bias=conv_bias)   #  lsnet/model/recattn.py:128-146)
        self._qk_cache = None
        self._pe_cache = None

    def _pe_params(self, device):
        """BatchNorm-folded `pe` filters [dim, 9] and bias [dim] (fp32); cached like the `qk` weights."""
        m = self.pe
        src = [m.conv.weight, m.conv.bias, m.norm.weight, m.norm.bias, m.norm.running_mean, m.norm.running_var] if isinstance(m, ConvNorm) else [m.weight, m.bias]
        key = (device,) + tuple((id(t), t._version) for t in src if t is not None)
        c = self._pe_cache
        if c is None or c[0] != key:
            w, b = _wb(m)
            c = (key, w.detach().float().reshape(w.shape[0], 9).contiguous(), None if b is None else b.detach().float().contiguous())
            self._pe_cache = c
        return c[1], c[2]

    def _qk_params(self, dtype, device):
        """BatchNorm-folded `qk` weights as [2, dim, dim / 2] in the activation dtype and the fp32 bias [2 dim]; cached, keyed on the version
        counters of the source tensors (load_state_dict / in-place updates invalidate it)."""
        m = self.qk
        src = [m.conv.weight, m.conv.bias, m.norm.weight, m.norm.bias, m.norm.running_mean, m.norm.running_var] if isinstance(m, ConvNorm) else [m.weight, m.bias]
        key = (dtype, device) + tuple((id(t), t._version) for t in src if t is not None)
        c = self._qk_cache
        if c is None or c[0] != key:
            w, b = _wb(m)
            dim = w.shape[0] // 2
            c = (key, w.detach().reshape(2, dim, dim // 2).to(dtype).contiguous(), b.detach().float().contiguous())
            self._qk_cache = c
        return c[1], c[2]

    def forward(self, x):
        if self.training:
            raise RuntimeError("recnext_b200 linear attention: the sm_100a kernel has no backward — call .eval(); there is deliberately no PyTorch fallback")
        if not x.is_cuda:
            raise RuntimeError("recnext_b200 linear attention runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.head_dim not in LINATTN_HEAD_DIMS:
            raise RuntimeError(f"recnext_b200 linear attention: head_dim {self.head_dim} is not built (have {LINATTN_HEAD_DIMS})")
        _no_grad_wanted(x, self, "linear attention")
        if torch.is_autocast_enabled():
            x = x.to(torch.get_autocast_dtype("cuda"))          # (what the convs of the library path do to their input under autocast)
        if not LINATTN_GEMM or x.dtype not in _DTYPES:
            return linattn_forward(self.qk(x), x, self.pe(x), self.num_heads)
        x = x.contiguous()
        B, dim, H, W = x.shape
        w, b = self._qk_params(x.dtype, x.device)
        xg = x.view(B, 2, dim // 2, H * W)
        q = torch.bmm(w[0].expand(B, dim, dim // 2), xg[:, 0])     # [B, dim, n]: batched GEMMs with the weight broadcast over the batch (stride 0)
        k = torch.bmm(w[1].expand(B, dim, dim // 2), xg[:, 1])
        pe_conv = self.pe.conv if isinstance(self.pe, ConvNorm) else self.pe
        if tuple(pe_conv.kernel_size) == (3, 3) and tuple(pe_conv.padding) == (1, 1) and tuple(pe_conv.stride) == (1, 1) and pe_conv.groups == dim:
            pw, pb = self._pe_params(x.device)
            return linattn_forward_pe(q, k, b[:dim], b[dim:], x, pw, pb, self.num_heads)    # `pe` evaluated inside the kernel
        return linattn_forward_qk(q, k, b[:dim], b[dim:], x, self.pe(x), self.num_heads)


class LinearAttention1(_LinearAttention):
    """model/recattn.py:8-28"""


class LinearAttention2(_LinearAttention):
    """model/recattn.py:31-51 (same function, model/recattn.py:56-57)"""


class RecAttn2d(nn.Module):
    """``RecAttn2d(dim, num_heads, kernel_size=5, stage=1, mode='nearest')`` — reference model/recattn.py:54-67."""

    def __init__(self, dim, num_heads, kernel_size=5, stage=1, mode="nearest", conv_bias=False):
        super().__init__()
        if kernel_size != 5:
            raise ValueError("RecAttn2d: the CUDA kernels are built for kernel_size 5 (the reference's value)")
        if mode not in _MODES:
            raise ValueError(f"RecAttn2d: mode {mode!r} not supported (bilinear, nearest)")
        self.mode = mode
        LinearAttention = LinearAttention2 if stage >= 3 else LinearAttention1
        self.down = nn.Sequential(
            ConvNorm(dim, dim, kernel_size=kernel_size, padding=kernel_size // 2, stride=2, groups=dim, bias=conv_bias),
            LinearAttention(dim=dim, num_heads=num_heads, conv_bias=conv_bias),
        )
        self.conv = ConvNorm(dim, dim, kernel_size=kernel_size, padding=kernel_size // 2, groups=dim, bias=conv_bias)

    def forward(self, x):
        if self.training:
            raise RuntimeError("recnext_b200.RecAttn2d: the sm_100a kernels have no backward yet — call .eval() "
                               "(inference with BatchNorm folded); there is deliberately no PyTorch fallback")
        _no_grad_wanted(x, self, "RecAttn2d")
        if torch.is_autocast_enabled() and x.is_cuda:
            x = x.to(torch.get_autocast_dtype("cuda"))
        wd, bd = _wb(self.down[0])
        wc, bc = _wb(self.conv)
        low = recattn_down_forward(x, wd, bd)                 # model/recattn.py:60
        z = self.down[1](low)                                 # linear attention: library GEMMs, as the reference writes it
        return recattn_up_forward(x, z, wc, bc, self.mode)    # model/recattn.py:67
