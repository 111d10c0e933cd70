"""The reference's other wirings of the hot path, served by the same sm_100a kernels.

* ``MllaRecConv2d`` — the RecConv2d of the MLLA ablation (reference mlla/models/mlla_recconv.py:20-50): the up path is
  ``nn.Upsample(scale_factor=2, mode)`` instead of ``F.interpolate(size=...)``, default mode 'nearest'.  That module only works
  when every level halves exactly (H, W divisible by 2**level: its own comment says "only support resolutions like 256, 384";
  otherwise ``f + x`` fails on a shape mismatch); there a scale-factor-2 upsample and an interpolate to the exact size use the same
  source indices and weights (scale 1/2 either way), so the fused RecConv kernels compute it unchanged.  Same parameters and
  ``state_dict`` keys (``down``, ``convs.{j}``; ``up`` has none).
* ``LsRecAttn2d`` / ``PartialChannelOperation`` — the L-series token mixer (reference lsnet/model/recattn.py:115-127, :226-237):
  RecAttn2d on the first ``1 / split_rate`` of the channels, the rest passed through; the linear attention is picked by
  ``[LinearAttention1, LinearAttention2, LinearAttention2][stage]`` (the two are the same function, :480-501).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .recattn import LinearAttention1, LinearAttention2, RecAttn2d
from .recconv import RecConv2d


class MllaRecConv2d(RecConv2d):
    """``RecConv2d(in_channels, kernel_size=5, bias=False, level=2, mode='nearest')`` of mlla/models/mlla_recconv.py:20-50."""

    def __init__(self, in_channels, kernel_size=5, bias=False, level=2, mode="nearest"):
        super().__init__(in_channels, kernel_size=kernel_size, bias=bias, level=level, mode=mode)
        self.up = nn.Upsample(scale_factor=2, mode=mode)   # parameter-free; kept so that the module tree prints like the reference's

    def forward(self, x):
        H, W = x.shape[-2:]
        m = 1 << self.level
        if H % m or W % m:
            # the reference fails in `f + x` (the upsampled map is larger than the feature it is added to)
            raise ValueError(f"MllaRecConv2d: H and W must be divisible by 2**level = {m} (got {H}x{W}); the scale-factor-2 up path of "
                             "mlla/models/mlla_recconv.py:37-50 has no other sizes")
        return super().forward(x)


class LsRecAttn2d(RecAttn2d):
    """``RecAttn2d(dim, num_heads, kernel_size=5, stage=1, mode='nearest')`` of lsnet/model/recattn.py:115-127."""

    def __init__(self, dim, num_heads, kernel_size=5, stage=1, mode="nearest"):
        super().__init__(dim, num_heads, kernel_size=kernel_size, stage=0, mode=mode, conv_bias=True)   # the L-series ConvNorm keeps the conv bias (:128-146)
        if [LinearAttention1, LinearAttention2, LinearAttention2][stage] is LinearAttention2:
            self.down[1] = LinearAttention2(dim=dim, num_heads=num_heads, conv_bias=True)


class PartialChannelOperation(nn.Module):
    """``attn`` on the first ``in_channels / split_rate`` channels, identity on the rest (lsnet/model/recattn.py:226-237)."""

    def __init__(self, in_channels, attn, split_rate=4):
        super().__init__()
        assert in_channels % split_rate == 0, "in_channels must be divisible by split_rate"
        self.split_idx = in_channels // split_rate
        self.attn = attn

    def forward(self, x):
        x1 = self.attn(x[:, :self.split_idx].contiguous())   # the kernels take NCHW-contiguous planes
        return torch.cat([x1, x[:, self.split_idx:].to(x1.dtype)], dim=1)
