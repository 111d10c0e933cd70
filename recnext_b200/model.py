"""RecNeXt M-series classifier — host-side mirror of the reference model definition around the RecConv hot path.

The callers of the hot path: ``MetaNeXtBlock`` = ``x + drop_path(mlp(BN(RecConv2d(x))))``
(reference model/recnext.py:149-158), stem / downsample / classifier (:134-201), the variants m0..m5 (:365-407)
and BN folding for the fused-BN eval model (ConvNorm.fuse :75-97, NormLinear.fuse :109-122,
RecNextClassifier.fuse :191-201, utils.replace_batchnorm utils.py:227-234).  Module names are kept so that the
``state_dict`` keys and shapes are identical to the reference's (released checkpoints load with strict=True);
everything except the token mixer is plain PyTorch (out of scope for the CUDA work).

``token_mixer`` lets a caller pick the RecConv2d implementation; the default is the CUDA one.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.nn as nn

import ctypes
import os

from . import _native as N
from .recattn import RecAttn2d, _timing_start, _timing_stop
from .recconv import _DTYPES, RecConv2d, _stream

# RECNEXT_FFN=0 keeps the channel mixer on the library path (cuDNN 1x1 convs) for A/B measurements
FUSED_FFN = os.environ.get("RECNEXT_FFN", "1") != "0"


# RECNEXT_FFN_IMPL=mma keeps the round-1 mma.sync channel-mixer kernel (A/B measurements); default: the tcgen05 kernel
FFN_IMPL = os.environ.get("RECNEXT_FFN_IMPL", "tc")


def ffn_pack(w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """Packs w1 [hidden, C] and w2 [C, hidden] (16-bit CUDA tensors) into the weight stream of the tcgen05 channel-mixer kernel:
    128 x 64 tiles in the kernel's consumption order, each stored as the shared-memory image of a K-major tcgen05 operand, so that
    a tile is one contiguous TMA bulk copy (``recnext_ffn_pack``)."""
    hid, C = w1.shape
    if not (w1.is_cuda and w2.is_cuda) or w1.dtype not in (torch.bfloat16, torch.float16) or w2.dtype != w1.dtype or tuple(w2.shape) != (C, hid):
        raise ValueError(f"ffn_pack: w1 {tuple(w1.shape)} {w1.dtype} / w2 {tuple(w2.shape)} {w2.dtype} must be 16-bit CUDA [hidden, C] / [C, hidden]")
    with torch.cuda.device(w1.device):
        nbytes = N.lib().recnext_ffn_packed_bytes(C, hid)
        if nbytes == 0:
            raise RuntimeError(f"ffn_pack: C={C} hidden={hid} is not served by the tcgen05 channel-mixer kernel (C % 8 == 0, C <= 768)")
        packed = torch.empty((nbytes,), dtype=torch.uint8, device=w1.device)
        N.check(N.lib().recnext_ffn_pack(C, hid, _DTYPES[w1.dtype], w1.contiguous().data_ptr(), w2.contiguous().data_ptr(), packed.data_ptr(),
                                         _stream(w1)), "recnext_ffn_pack")
    return packed


def ffn_forward_packed(y: torch.Tensor, x: torch.Tensor, packed: torch.Tensor, b1: torch.Tensor, b2: torch.Tensor, hid: int) -> torch.Tensor:
    """out = x + b2 + W2 gelu(W1 y + b1) with (W1, W2) given as ``ffn_pack(w1, w2)`` (``recnext_ffn_forward_packed``)."""
    if not (y.is_cuda and x.is_cuda):
        raise RuntimeError("recnext_b200.ffn_forward runs on CUDA (sm_100a) only; there is no CPU fallback")
    if y.dtype not in (torch.bfloat16, torch.float16) or x.dtype != y.dtype:
        raise TypeError("ffn_forward: y and x must share a 16-bit dtype (bfloat16 / float16)")
    y, x = y.contiguous(), x.contiguous()
    B, C, H, W = y.shape
    if tuple(x.shape) != tuple(y.shape) or b1.numel() != hid or b2.numel() != C:
        raise ValueError(f"ffn_forward: shapes y {tuple(y.shape)} x {tuple(x.shape)} b1 {tuple(b1.shape)} b2 {tuple(b2.shape)} hidden {hid}")
    out = torch.empty_like(y)
    with torch.cuda.device(y.device):
        ev = _timing_start()
        N.check(N.lib().recnext_ffn_forward_packed(B, C, hid, H * W, _DTYPES[y.dtype], y.data_ptr(), x.data_ptr(), packed.data_ptr(),
                                                   b1.data_ptr(), b2.data_ptr(), out.data_ptr(), _stream(y)), "recnext_ffn_forward_packed")
        _timing_stop(ev, 3 * y.numel() * y.element_size(), ("ffn",) + tuple(y.shape))
    return out


def ffn_forward(y: torch.Tensor, x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """out = x + b2 + W2 gelu(W1 y + b1) per pixel on NCHW tensors — the channel mixer + residual of a RecNeXt block
    (reference model/recnext.py:125-131,157-158 with every BatchNorm folded) as ONE sm_100a kernel on the tcgen05 tensor
    cores (``recnext_ffn_pack`` + ``recnext_ffn_forward_packed``, include/recnext_b200.h).  16-bit CUDA tensors; no fallback:
    unsupported shapes raise."""
    if not (y.is_cuda and x.is_cuda):
        raise RuntimeError("recnext_b200.ffn_forward runs on CUDA (sm_100a) only; there is no CPU fallback")
    if y.dtype not in (torch.bfloat16, torch.float16) or x.dtype != y.dtype or w1.dtype != y.dtype or w2.dtype != y.dtype:
        raise TypeError("ffn_forward: y, x, w1, w2 must share a 16-bit dtype (bfloat16 / float16)")
    B, C, H, W = y.shape
    hid = w1.shape[0]
    if tuple(w1.shape) != (hid, C) or tuple(w2.shape) != (C, hid) or tuple(x.shape) != tuple(y.shape):
        raise ValueError(f"ffn_forward: shapes y {tuple(y.shape)} x {tuple(x.shape)} w1 {tuple(w1.shape)} w2 {tuple(w2.shape)}")
    if FFN_IMPL == "mma":
        y, x = y.contiguous(), x.contiguous()
        out = torch.empty_like(y)
        with torch.cuda.device(y.device):
            ev = _timing_start()
            N.check(N.lib().recnext_ffn_forward(B, C, hid, H * W, _DTYPES[y.dtype], y.data_ptr(), x.data_ptr(), w1.contiguous().data_ptr(),
                                                b1.float().contiguous().data_ptr(), w2.contiguous().data_ptr(), b2.float().contiguous().data_ptr(),
                                                out.data_ptr(), _stream(y)), "recnext_ffn_forward")
            _timing_stop(ev, 3 * y.numel() * y.element_size(), ("ffn",) + tuple(y.shape))
        return out
    return ffn_forward_packed(y, x, ffn_pack(w1, w2), b1.float().contiguous(), b2.float().contiguous(), hid)


def _param_key(*mods) -> tuple:
    """Identity + version of every parameter / buffer the folded weights are derived from: load_state_dict(), an EMA copy_ or any
    other in-place update bumps ``_version`` and invalidates the cache."""
    key = []
    for m in mods:
        if m is None:
            continue
        for t in list(m.parameters(recurse=False)) + list(m.buffers(recurse=False)):
            key.append((t.data_ptr(), t._version))
    return tuple(key)


def _needs_autograd(*tensors_and_modules) -> bool:
    """The fused inference kernels run outside autograd: they are only taken when no gradient can be asked for."""
    if not torch.is_grad_enabled():
        return False
    for o in tensors_and_modules:
        if isinstance(o, torch.Tensor):
            if o.requires_grad:
                return True
        elif o is not None and any(p.requires_grad for p in o.parameters()):
            return True
    return False


VARIANTS = {  # reference model/recnext.py:369-406
    "recnext_m0": dict(embed_dim=(40, 80, 160, 320), depth=(2, 2, 9, 1)),
    "recnext_m1": dict(embed_dim=(48, 96, 192, 384), depth=(3, 3, 15, 2)),
    "recnext_m2": dict(embed_dim=(56, 112, 224, 448), depth=(3, 3, 15, 2)),
    "recnext_m3": dict(embed_dim=(64, 128, 256, 512), depth=(3, 3, 13, 2)),
    "recnext_m4": dict(embed_dim=(64, 128, 256, 512), depth=(5, 5, 25, 4)),
    "recnext_m5": dict(embed_dim=(80, 160, 320, 640), depth=(7, 7, 35, 2)),
    # A-series (linear attention + nearest interpolation), reference model/recattn.py:380-426
    "recnext_a0": dict(series="a", embed_dim=(40, 80, 160, 320), depth=(2, 2, 9, 1)),
    "recnext_a1": dict(series="a", embed_dim=(48, 96, 192, 384), depth=(3, 3, 15, 2)),
    "recnext_a2": dict(series="a", embed_dim=(56, 112, 224, 448), depth=(3, 3, 15, 2)),
    "recnext_a3": dict(series="a", embed_dim=(64, 128, 256, 512), depth=(3, 3, 13, 2), mlp_ratio=1.875),
    "recnext_a4": dict(series="a", embed_dim=(64, 128, 256, 512), depth=(5, 5, 25, 4), mlp_ratio=1.875),
    "recnext_a5": dict(series="a", embed_dim=(80, 160, 320, 640), depth=(7, 7, 35, 2), mlp_ratio=1.875),
}


class DropPath(nn.Module):
    """Stochastic depth per sample (what timm.layers.DropPath does in training mode)."""

    def __init__(self, p: float = 0.0):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1.0 - self.p
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask.div_(keep)


class ConvNorm(nn.Sequential):
    """conv (no bias) + BatchNorm2d; ``fuse()`` folds the running statistics into a biased conv."""

    def __init__(self, cin, cout, kernel_size=1, stride=1, padding=0, groups=1):
        super().__init__()
        self.add_module("conv", nn.Conv2d(cin, cout, kernel_size, stride, padding, groups=groups, bias=False))
        self.add_module("norm", nn.BatchNorm2d(cout))

    @torch.no_grad()
    def fuse(self) -> nn.Conv2d:
        conv, bn = self.conv, self.norm
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shift = bn.bias - scale * bn.running_mean
        if conv.bias is not None:
            shift = shift + scale * conv.bias
        out = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation,
                  conv.groups, bias=True, device=conv.weight.device, dtype=conv.weight.dtype)
        out.weight.copy_(conv.weight * scale.view(-1, 1, 1, 1))
        out.bias.copy_(shift)
        return out


class NormLinear(nn.Sequential):
    """BatchNorm1d + Linear; ``fuse()`` folds the norm into the linear layer."""

    def __init__(self, cin, cout, std=0.02):
        super().__init__()
        self.add_module("norm", nn.BatchNorm1d(cin))
        self.add_module("linear", nn.Linear(cin, cout, bias=True))
        nn.init.trunc_normal_(self.linear.weight, std=std)
        nn.init.zeros_(self.linear.bias)

    @torch.no_grad()
    def fuse(self) -> nn.Linear:
        bn, lin = self.norm, self.linear
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        out = nn.Linear(lin.in_features, lin.out_features, device=lin.weight.device, dtype=lin.weight.dtype)
        out.weight.copy_(lin.weight * scale.view(1, -1))
        out.bias.copy_(lin.weight @ shift + lin.bias)
        return out


def mlp(cin: int, hidden: int, act_layer=nn.GELU) -> nn.Sequential:
    hidden = int(hidden)
    return nn.Sequential(ConvNorm(cin, hidden, 1), act_layer(), ConvNorm(hidden, cin, 1))


FUSED_STEM = os.environ.get("RECNEXT_STEM", "1") != "0"


def stem_pack(conv1: nn.Conv2d, conv2: nn.Conv2d, dtype):
    """Weights of the fused stem kernel from the two (BatchNorm-folded) 3x3 stride-2 convs:
    w1p [C1P, 32] with k = ci * 9 + ky * 3 + kx, w2p [C2P, 9 * C1P] with k = (ky * 3 + kx) * C1P + ci, zero padded; fp32 biases."""
    with torch.no_grad():
        C1, C2 = conv1.out_channels, conv2.out_channels
        C1P, C2P = (32 if C1 <= 32 else 48), (C2 + 15) // 16 * 16
        dev = conv1.weight.device
        w1p = torch.zeros(C1P, 32, device=dev, dtype=torch.float32)
        w1p[:C1, :27] = conv1.weight.float().reshape(C1, 27)
        w2p = torch.zeros(C2P, 9, C1P, device=dev, dtype=torch.float32)
        w2p[:C2, :, :C1] = conv2.weight.float().permute(0, 2, 3, 1).reshape(C2, 9, C1)
        b1p = torch.zeros(C1P, device=dev, dtype=torch.float32)
        b1p[:C1] = conv1.bias.float()
        b2p = torch.zeros(C2P, device=dev, dtype=torch.float32)
        b2p[:C2] = conv2.bias.float()
        return w1p.to(dtype).contiguous(), b1p, w2p.reshape(C2P, 9 * C1P).to(dtype).contiguous(), b2p, C1, C2


def stem_forward(x: torch.Tensor, w1p, b1p, w2p, b2p, C1: int, C2: int) -> torch.Tensor:
    """conv 3x3 s2 -> GELU -> conv 3x3 s2 of the stem (reference model/recnext.py:139-146, ConvNorms folded) as ONE sm_100a kernel
    (``recnext_stem_forward``): the 112 x 112 intermediate never leaves shared memory.  x: [B, 3, H, W] 16-bit CUDA tensor."""
    if not x.is_cuda:
        raise RuntimeError("recnext_b200.stem_forward runs on CUDA (sm_100a) only; there is no CPU fallback")
    if x.dtype not in (torch.bfloat16, torch.float16) or x.dim() != 4 or x.shape[1] != 3:
        raise TypeError("stem_forward: x must be a 16-bit [B, 3, H, W] tensor")
    x = x.contiguous()
    B, _, H, W = x.shape
    H2, W2 = ((H - 1) // 2) // 2 + 1, ((W - 1) // 2) // 2 + 1
    out = torch.empty(B, C2, H2, W2, device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        ev = _timing_start()
        N.check(N.lib().recnext_stem_forward(B, H, W, C1, C2, _DTYPES[x.dtype], x.data_ptr(), w1p.data_ptr(), b1p.data_ptr(), w2p.data_ptr(),
                                             b2p.data_ptr(), out.data_ptr(), _stream(x)), "recnext_stem_forward")
        _timing_stop(ev, (x.numel() + out.numel()) * x.element_size(), ("stem",) + tuple(x.shape))
    return out


class RecNextStem(nn.Module):
    def __init__(self, cin, cout, act_layer=nn.GELU):
        super().__init__()
        self.stem = nn.Sequential(ConvNorm(cin, cout // 2, 3, 2, 1), act_layer(), ConvNorm(cout // 2, cout, 3, 2, 1))

    def train(self, mode: bool = True):
        self._stem_cache = None
        return super().train(mode)

    def _fused_ok(self, x) -> bool:
        """Eval mode, ConvNorms folded (fuse()), 16-bit compute, no gradient wanted: the fused kernel takes the stem."""
        if self.training or not x.is_cuda or not FUSED_STEM or torch.jit.is_tracing() or x.dim() != 4 or x.shape[1] != 3:
            return False
        c1, act, c2 = self.stem[0], self.stem[1], self.stem[2]
        if not (type(c1) is nn.Conv2d and type(c2) is nn.Conv2d and c1.bias is not None and c2.bias is not None):
            return False
        if not isinstance(act, nn.GELU) or getattr(act, "approximate", "none") != "none":
            return False
        if c1.out_channels > 48 or c2.out_channels > 80 or c1.in_channels != 3:
            return False
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        return dt in (torch.bfloat16, torch.float16) and not _needs_autograd(x, self)

    def forward(self, x):
        if self._fused_ok(x):
            dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
            key = (dt, x.device, _param_key(self.stem[0], self.stem[2]))
            c = getattr(self, "_stem_cache", None)
            if c is None or c[0] != key:
                c = (key, stem_pack(self.stem[0], self.stem[2], dt))
                self._stem_cache = c
            return stem_forward(x.to(dt), *c[1])
        return self.stem(x)


def dwdown_forward(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """norm(token_mixer(x)) of a Downsample block: depthwise 7x7 stride-2 conv, channel multiplier 2, BatchNorm folded into (w, b)
    (reference model/recnext.py:137-138,145) as one sm_100a kernel (``recnext_dwdown_forward``).  CUDA tensors, fp32 or 16-bit."""
    if not x.is_cuda:
        raise RuntimeError("recnext_b200.dwdown_forward runs on CUDA (sm_100a) only; there is no CPU fallback")
    if x.dtype not in _DTYPES:
        raise TypeError("dwdown_forward: float32 / bfloat16 / float16 only")
    x = x.contiguous()
    B, C, H, W = x.shape
    if tuple(w.shape) != (2 * C, 1, 7, 7) or tuple(b.shape) != (2 * C,):
        raise ValueError(f"dwdown_forward: w {tuple(w.shape)} / b {tuple(b.shape)} do not match x {tuple(x.shape)}")
    out = torch.empty(B, 2 * C, (H - 1) // 2 + 1, (W - 1) // 2 + 1, device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        ev = _timing_start()
        N.check(N.lib().recnext_dwdown_forward(B, C, H, W, _DTYPES[x.dtype], x.data_ptr(), w.float().contiguous().data_ptr(),
                                               b.float().contiguous().data_ptr(), out.data_ptr(), _stream(x)), "recnext_dwdown_forward")
        _timing_stop(ev, (x.numel() + out.numel()) * x.element_size(), ("dwdown",) + tuple(x.shape))
    return out


def _fold_mlp(fc1: nn.Conv2d, fc2: nn.Conv2d, dtype, bn: Optional[nn.BatchNorm2d] = None):
    """(w1, b1, w2, b2) for ffn_forward from the two folded 1x1 convs of an `mlp`; an eval-mode BatchNorm in front of it is folded
    into (w1, b1): W1 (s*y + t) + b1 = (W1 diag(s)) y + (b1 + W1 t).  The hidden width is zero-padded to a multiple of 32
    (gelu(0) = 0 meets zero columns of W2), so e.g. the A-series' 1.875 ratio (120 hidden channels at C = 64) is served too."""
    with torch.no_grad():
        w1 = fc1.weight.float().view(fc1.out_channels, fc1.in_channels)
        w2 = fc2.weight.float().view(fc2.out_channels, fc2.in_channels)
        b1, b2 = fc1.bias.float(), fc2.bias.float()
        if bn is not None:
            s = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float()
            t = (bn.bias - s * bn.running_mean).float()
            b1 = b1 + w1 @ t
            w1 = w1 * s.view(1, -1)
        hid = w1.shape[0]
        pad = (-hid) % 32 if FFN_IMPL == "mma" else 0   # (the tcgen05 kernel pads the hidden width to 128-row chunks itself)
        if pad:
            w1 = torch.cat([w1, w1.new_zeros(pad, w1.shape[1])], 0)
            b1 = torch.cat([b1, b1.new_zeros(pad)], 0)
            w2 = torch.cat([w2, w2.new_zeros(w2.shape[0], pad)], 1)
        return w1.to(dtype).contiguous(), b1.contiguous(), w2.to(dtype).contiguous(), b2.contiguous()


def _prepare_ffn(w1, b1, w2, b2):
    """Folded weights -> what the kernel of the selected implementation consumes (the tcgen05 kernel: the packed weight stream)."""
    if FFN_IMPL == "mma":
        return ("mma", w1, b1, w2, b2)
    return ("tc", ffn_pack(w1, w2), b1, b2, w1.shape[0])


def _ffn_apply(y, x, prepared):
    if prepared[0] == "mma":
        return ffn_forward(y, x, *prepared[1:])
    return ffn_forward_packed(y, x, *prepared[1:])


def _ffn_shape_ok(block: nn.Module, x: torch.Tensor) -> bool:
    """Does this block (eval mode, ConvNorms folded, no gradient wanted) take the fused channel-mixer kernel for input x?"""
    if block.training or not x.is_cuda or not FUSED_FFN or torch.jit.is_tracing():
        return False
    fc1, fc2 = block.channel_mixer[0], block.channel_mixer[2]
    if not (type(fc1) is nn.Conv2d and type(fc2) is nn.Conv2d and fc1.bias is not None and fc2.bias is not None):
        return False  # ConvNorms not folded yet (replace_batchnorm / fuse())
    if not isinstance(block.channel_mixer[1], nn.GELU) or getattr(block.channel_mixer[1], "approximate", "none") != "none":
        return False
    if _needs_autograd(x, block):
        return False  # a gradient may be asked for (fine-tuning with frozen BN, saliency): the differentiable op chain runs instead
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
    if x.dtype != dt and isinstance(block, Downsample):
        return False
    C, HW = fc1.in_channels, x.shape[2] * x.shape[3]
    if dt not in (torch.bfloat16, torch.float16):
        return False
    if FFN_IMPL == "mma":
        hid = (fc1.out_channels + 31) // 32 * 32
        staged = 192 <= C <= 256 and C % 32 == 0 and hid % 32 == 0
        return (C % 16 == 0 and hid % 16 == 0 and HW % 4 == 0 and (HW >= 400 or staged or os.environ.get("RECNEXT_FFN") == "all")
                and (2 * C + hid) * 144 + 64 <= 227 * 1024)
    # tcgen05 kernel: every RecNeXt width (40 .. 640, hidden 2C or 1.875C) and every plane size
    return C % 8 == 0 and C <= 768


class MetaNeXtBlock(nn.Module):
    """x + drop_path(channel_mixer(norm(token_mixer(x)))), token_mixer = RecConv2d(level = 4 - stage, k = 5)."""

    def __init__(self, dim, mlp_ratio, act_layer=nn.GELU, stage=0, drop_path=0.0, token_mixer: Callable = RecConv2d):
        super().__init__()
        self.token_mixer = token_mixer(dim, level=4 - stage, kernel_size=5)
        self.norm = nn.BatchNorm2d(dim)
        self.channel_mixer = mlp(dim, dim * mlp_ratio, act_layer)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def train(self, mode: bool = True):
        self._ffn_cache = None  # folded weights belong to the eval-mode statistics
        return super().train(mode)

    def _ffn_params(self, dtype, device):
        """Folded (and packed) weights of the fused channel mixer, eval-mode BatchNorm `norm` folded in.  Cached; the key holds the
        version counter of every source tensor, so load_state_dict() / in-place updates are picked up."""
        key = (dtype, device, _param_key(self.channel_mixer[0], self.channel_mixer[2], self.norm))
        c = getattr(self, "_ffn_cache", None)
        if c is None or c[0] != key:
            c = (key, _prepare_ffn(*_fold_mlp(self.channel_mixer[0], self.channel_mixer[2], dtype, self.norm)))
            self._ffn_cache = c
        return c[1]

    def _ffn_eligible(self, x) -> bool:
        return _ffn_shape_ok(self, x)

    def forward(self, x):
        if self._ffn_eligible(x):
            # token mixer (sm_100a RecConv kernel) -> ONE fused kernel for norm + 1x1 conv + GELU + 1x1 conv + residual
            y = self.token_mixer(x)
            return _ffn_apply(y, x.to(y.dtype), self._ffn_params(y.dtype, y.device))
        return x + self.drop_path(self.channel_mixer(self.norm(self.token_mixer(x))))


class MetaNeXtBlockA(nn.Module):
    """A-series block: x + drop_path(channel_mixer(token_mixer(x))), token_mixer = RecAttn2d(dim, heads = 2^(stage+1))
    (reference model/recattn.py:162-171; no BatchNorm between the mixers)."""

    def __init__(self, dim, mlp_ratio, act_layer=nn.GELU, stage=0, drop_path=0.0, token_mixer: Callable = RecAttn2d):
        super().__init__()
        self.token_mixer = token_mixer(dim, num_heads=2 ** (stage + 1), stage=stage)
        self.channel_mixer = mlp(dim, dim * mlp_ratio, act_layer)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def train(self, mode: bool = True):
        self._ffn_cache = None
        return super().train(mode)

    def forward(self, x):
        if _ffn_shape_ok(self, x):
            y = self.token_mixer(x)
            key = (y.dtype, y.device, _param_key(self.channel_mixer[0], self.channel_mixer[2]))
            c = getattr(self, "_ffn_cache", None)
            if c is None or c[0] != key:
                c = (key, _prepare_ffn(*_fold_mlp(self.channel_mixer[0], self.channel_mixer[2], y.dtype)))
                self._ffn_cache = c
            return _ffn_apply(y, x.to(y.dtype), c[1])
        return x + self.drop_path(self.channel_mixer(self.token_mixer(x)))


class Downsample(nn.Module):
    def __init__(self, dim, mlp_ratio, act_layer=nn.GELU):
        super().__init__()
        self.token_mixer = nn.Conv2d(dim, dim * 2, kernel_size=7, padding=3, groups=dim, stride=2)
        self.norm = nn.BatchNorm2d(dim * 2)
        self.channel_mixer = mlp(dim * 2, dim * 2 * mlp_ratio, act_layer)

    def train(self, mode: bool = True):
        self._ffn_cache = None
        self._dw_cache = None
        return super().train(mode)

    def _dw_params(self, device):
        key = (device, _param_key(self.token_mixer, self.norm))
        c = getattr(self, "_dw_cache", None)
        if c is None or c[0] != key:
            conv, bn = self.token_mixer, self.norm
            with torch.no_grad():
                s = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float()
                t = (bn.bias - s * bn.running_mean).float()
                w = (conv.weight.float() * s.view(-1, 1, 1, 1)).contiguous()
                b = (s * conv.bias.float() + t).contiguous() if conv.bias is not None else t.contiguous()
            c = (key, (w, b))
            self._dw_cache = c
        return c[1]

    def forward(self, x):
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        if (not self.training and x.is_cuda and FUSED_FFN and dt in (torch.bfloat16, torch.float16) and not torch.jit.is_tracing()
                and not _needs_autograd(x, self) and (x.shape[2] + 8) * ((x.shape[3] + 7) // 8 + 10) * 32 + 448 <= 227 * 1024):
            x = dwdown_forward(x.to(dt), *self._dw_params(x.device))   # norm(token_mixer(x)): one kernel, BatchNorm folded
        else:
            x = self.norm(self.token_mixer(x))
        if _ffn_shape_ok(self, x):
            # x + mlp(x) (model/recnext.py:145-146) as the same fused kernel: the mixer input is also the residual
            key = (x.dtype, x.device, _param_key(self.channel_mixer[0], self.channel_mixer[2]))
            c = getattr(self, "_ffn_cache", None)
            if c is None or c[0] != key:
                c = (key, _prepare_ffn(*_fold_mlp(self.channel_mixer[0], self.channel_mixer[2], x.dtype)))
                self._ffn_cache = c
            return _ffn_apply(x, x, c[1])
        return x + self.channel_mixer(x)


class RecNextClassifier(nn.Module):
    def __init__(self, dim, num_classes, distillation=False):
        super().__init__()
        self.num_classes, self.distillation = num_classes, distillation
        self.head_drop = nn.Dropout(0.0)
        self.head = NormLinear(dim, num_classes) if num_classes > 0 else nn.Identity()
        self.head_dist = NormLinear(dim, num_classes) if num_classes > 0 else nn.Identity()

    def forward(self, x):
        x = self.head_drop(x)
        a, b = self.head(x), self.head_dist(x)
        if self.training and self.distillation:
            return a, b
        return (a + b) / 2

    @torch.no_grad()
    def fuse(self):
        if not self.num_classes > 0:
            return nn.Identity()
        a, b = self.head.fuse(), self.head_dist.fuse()
        a.weight.copy_((a.weight + b.weight) / 2)
        a.bias.copy_((a.bias + b.bias) / 2)
        return a


class RecNextStage(nn.Module):
    def __init__(self, cin, cout, depth, mlp_ratio, act_layer, downsample, stage, drop_path, token_mixer, series="m"):
        super().__init__()
        self.downsample = Downsample(cin, mlp_ratio, act_layer) if downsample else nn.Identity()
        block = MetaNeXtBlockA if series == "a" else MetaNeXtBlock
        self.blocks = nn.Sequential(*[block(cout, mlp_ratio, act_layer, stage, drop_path, token_mixer) for _ in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class RecNext(nn.Module):
    def __init__(self, in_chans=3, embed_dim: Sequence[int] = (48,), depth: Sequence[int] = (2,), mlp_ratio=2, num_classes=1000,
                 act_layer=nn.GELU, distillation=False, drop_rate=0.0, drop_path=0.0, token_mixer: Callable = RecConv2d, series="m"):
        super().__init__()
        self.series = series
        self.embed_dim, self.num_classes, self.num_features = tuple(embed_dim), num_classes, embed_dim[-1]
        self.stem = RecNextStem(in_chans, embed_dim[0], act_layer)
        stages, cin = [], embed_dim[0]
        for i, (dim, d) in enumerate(zip(embed_dim, depth)):
            stages.append(RecNextStage(cin, dim, d, mlp_ratio, act_layer, i != 0, i, drop_path, token_mixer, series))
            cin = dim
        self.stages = nn.Sequential(*stages)
        self.head_drop = nn.Dropout(drop_rate)
        self.head = RecNextClassifier(embed_dim[-1], num_classes, distillation)

    def forward_features(self, x):
        return self.stages(self.stem(x))

    def forward_head(self, x):
        return self.head(self.head_drop(x.mean((2, 3))))

    def forward(self, x):
        return self.forward_head(self.forward_features(x))

    @torch.no_grad()
    def fuse(self):
        replace_batchnorm(self)


def replace_batchnorm(net: nn.Module) -> nn.Module:
    """Recursively swaps every child that has ``fuse()`` for its fused form (reference utils.py:227-234).
    As in the reference, MetaNeXtBlock.norm / Downsample.norm have no ``fuse`` and stay as BatchNorm2d."""
    for name, child in list(net.named_children()):
        if hasattr(child, "fuse"):
            fused = child.fuse()
            setattr(net, name, fused)
            replace_batchnorm(fused)
        else:
            replace_batchnorm(child)
    return net


def create_model(variant: str, token_mixer: Callable = None, **kwargs) -> RecNext:
    """``create_model('recnext_m3')`` / ``create_model('recnext_a3')`` — the timm entry points the reference registers
    (model/recnext.py:365-407, model/recattn.py:380-426)."""
    if variant not in VARIANTS:
        raise ValueError(f"unknown variant {variant!r}; available: {sorted(VARIANTS)}")
    args = dict(VARIANTS[variant])
    if variant[-2:] in ("m4", "m5", "a4", "a5") and not kwargs.get("distillation", False):
        args["drop_path"] = 0.2 if variant.endswith("4") else 0.3
    args.update(kwargs)
    if token_mixer is None:
        token_mixer = RecAttn2d if args.get("series", "m") == "a" else RecConv2d
    return RecNext(token_mixer=token_mixer, **args)


def swap_recconv(net: nn.Module, token_mixer: Callable = RecConv2d) -> nn.Module:
    """Replaces every module that looks like the reference RecConv2d (has ``down``, ``convs``, ``level``, ``mode``) by
    ``token_mixer`` carrying the same parameters — for models built from the reference's own code."""
    for name, child in list(net.named_children()):
        if all(hasattr(child, a) for a in ("down", "convs", "level", "mode")) and not isinstance(child, token_mixer):
            k = child.down.kernel_size[0]
            new = token_mixer(child.down.in_channels, kernel_size=k, bias=child.down.bias is not None, level=child.level,
                              mode=child.mode)
            new.load_state_dict(child.state_dict(), strict=True)
            new.to(child.down.weight.device, child.down.weight.dtype)
            setattr(net, name, new)
        else:
            swap_recconv(child, token_mixer)
    return net
