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
