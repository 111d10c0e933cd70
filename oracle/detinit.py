"""Deterministic, name-keyed parameter fill shared by the golden-vector generator and the tests.

TEST INFRASTRUCTURE.  Full-model checkpoints are too large to commit (M0 = 11 MB), so instead of
storing weights we regenerate them: every tensor of a state_dict is filled from a torch.Generator
seeded by (seed, crc32(name)), which makes the values independent of module construction order.
Scales keep activations O(1) through ~60 layers so logits stay informative.
"""
import zlib

import torch


@torch.no_grad()
def fill_state_dict_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    sd = module.state_dict()
    for name in sorted(sd):
        t = sd[name]
        if name.endswith("num_batches_tracked"):
            continue
        g = torch.Generator().manual_seed((seed << 32) ^ zlib.crc32(name.encode()))
        shape = tuple(t.shape)
        if name.endswith("running_var"):
            v = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif name.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif t.ndim == 1 and name.endswith("weight"):  # BN scale
            v = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif t.ndim == 1:  # any bias
            v = torch.randn(shape, generator=g) * 0.05
        else:  # conv / linear weight: unit-gain in fan_in
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            v = torch.randn(shape, generator=g) * fan_in ** -0.5
        t.copy_(v.to(t.dtype))
    return module
