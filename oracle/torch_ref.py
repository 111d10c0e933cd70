"""PyTorch-eager restatement of the reference RecConv2d — TEST / BASELINE INFRASTRUCTURE ONLY.

This is the "reference CPU path" that bench.py times on the GPU box's host cores (the reference itself is pure
PyTorch and cannot travel to the box) and a second checker for tests at sizes where the C oracle is slow.  It
restates reference model/recnext.py:8-34 with the same ATen ops (F.conv2d groups=C, F.interpolate(size=...)),
so on CPU it executes exactly the kernels the reference would.  Pinned against tests/golden/ (generated from
the unmodified reference) in tests/test_torch_ref.py.  The product package never imports this module.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def recconv_reference(x: torch.Tensor, down_w: torch.Tensor, convs_w: List[torch.Tensor], down_b: Optional[torch.Tensor] = None,
                      convs_b: Optional[List[torch.Tensor]] = None, mode: str = "bilinear") -> torch.Tensor:
    """down pass (model/recnext.py:27-29), up pass (:31-33), final conv (:34)."""
    C = x.shape[1]
    k = down_w.shape[-1]
    L = len(convs_w) - 1
    bias = (lambda j: convs_b[j]) if convs_b is not None else (lambda j: None)
    feats = []
    cur = x
    for _ in range(L):
        size = cur.shape[2:]
        cur = F.conv2d(cur, down_w, down_b, stride=2, padding=k // 2, groups=C)
        feats.append((cur, size))
    up = 0
    for j, (f, size) in enumerate(reversed(feats)):
        up = F.interpolate(F.conv2d(f + up, convs_w[j], bias(j), padding=k // 2, groups=C), size=size, mode=mode)
    return F.conv2d(x + up, convs_w[L], bias(L), padding=k // 2, groups=C)


class RefRecConv2d(nn.Module):
    """Same constructor / state_dict as the reference module; forward = recconv_reference (ATen ops)."""

    def __init__(self, in_channels, kernel_size=5, bias=False, level=2, mode="bilinear"):
        super().__init__()
        self.level, self.mode = level, mode
        kw = dict(in_channels=in_channels, out_channels=in_channels, groups=in_channels, kernel_size=kernel_size,
                  padding=kernel_size // 2, bias=bias)
        self.down = nn.Conv2d(stride=2, **kw)
        self.convs = nn.ModuleList([nn.Conv2d(**kw) for _ in range(level + 1)])

    def forward(self, x):
        has_b = self.down.bias is not None
        return recconv_reference(x, self.down.weight, [c.weight for c in self.convs], self.down.bias if has_b else None,
                                 [c.bias for c in self.convs] if has_b else None, self.mode)


# ---------------------------------------------------------------------------------------------------------
# A-series token mixer (reference model/recattn.py:8-111) restated with the same ATen ops — the CPU baseline of
# bench.py --model recnext_a3 and a second checker for tests.  Pinned against tests/golden/recattn_*.npz.
# ---------------------------------------------------------------------------------------------------------
class RefConvNorm(nn.Sequential):
    def __init__(self, cin, cout, kernel_size=1, stride=1, padding=0, groups=1):
        super().__init__()
        self.add_module("conv", nn.Conv2d(cin, cout, kernel_size, stride, padding, 1, groups, bias=False))
        self.add_module("norm", nn.BatchNorm2d(cout))

    @torch.no_grad()
    def fuse(self):
        w = self.norm.weight / (self.norm.running_var + self.norm.eps) ** 0.5
        b = self.norm.bias - w * self.norm.running_mean
        w = w[:, None, None, None] * self.conv.weight
        c = self.conv
        m = nn.Conv2d(w.size(1) * c.groups, w.size(0), w.shape[2:], stride=c.stride, padding=c.padding, groups=c.groups)
        m.weight.data.copy_(w)
        m.bias.data.copy_(b)
        return m


class RefLinearAttention(nn.Module):
    def __init__(self, dim, num_heads, quadratic):
        super().__init__()
        self.num_heads, self.head_dim, self.quadratic = num_heads, dim // num_heads, quadratic
        self.qk = RefConvNorm(dim, dim * 2, 1, groups=2)
        self.pe = RefConvNorm(dim, dim, 3, padding=1, groups=dim)

    def forward(self, x):
        b, c, h, w = x.shape
        n = h * w
        s = n ** -0.5
        qk = F.elu(self.qk(x)) + 1.0
        (q, k), v = qk.view(b, 2, self.num_heads, self.head_dim, n).unbind(dim=1), x
        vt = v.view(b, self.num_heads, self.head_dim, n).transpose(-1, -2) * s
        if self.quadratic:  # LinearAttention2, model/recattn.py:39-51
            a = q.transpose(-1, -2) @ k
            a = a / (a.mean(dim=-1, keepdim=True) + 1e-6)
            out = (a * s) @ vt
        else:               # LinearAttention1, model/recattn.py:16-28
            q_t = q.transpose(-1, -2)
            out = q_t @ ((k * s) @ vt) / (q_t @ k.mean(dim=-1, keepdim=True) + 1e-6)
        return out.transpose(-1, -2).reshape(b, c, h, w) + self.pe(v)


class RefRecAttn2d(nn.Module):
    """Same constructor / state_dict as the reference RecAttn2d (model/recattn.py:54-67); forward = ATen ops."""

    def __init__(self, dim, num_heads, kernel_size=5, stage=1, mode="nearest"):
        super().__init__()
        self.mode = mode
        self.down = nn.Sequential(RefConvNorm(dim, dim, kernel_size, 2, kernel_size // 2, dim), RefLinearAttention(dim, num_heads, stage >= 3))
        self.conv = RefConvNorm(dim, dim, kernel_size, 1, kernel_size // 2, dim)

    def forward(self, x):
        return self.conv(x + F.interpolate(self.down(x), size=x.shape[2:], mode=self.mode))
