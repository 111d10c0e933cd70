"""Generates tests/golden/variant_*.npz by running the UNMODIFIED reference variants of the hot path in this container
(through oracle/timm_shim; timm is third-party and absent).  TEST INFRASTRUCTURE.

    python oracle/gen_golden_variants.py

* MLLA ablation RecConv2d (mlla/models/mlla_recconv.py:20-50): the up path is nn.Upsample(scale_factor=2, mode) instead of
  F.interpolate(size=...), default mode 'nearest';
* L-series token mixer (lsnet/model/recattn.py:115-127 RecAttn2d, :226-237 PartialChannelOperation): RecAttn2d on the first
  quarter of the channels, the rest passed through.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
import timm.layers as _layers  # noqa: E402  (shim)

# `timm.models.layers` (the old import path both files use) -> the shim's layers
_layers.to_2tuple = getattr(_layers, "to_2tuple", lambda v: v if isinstance(v, tuple) else (v, v))
pkg = types.ModuleType("timm.models.layers")
pkg.__dict__.update({k: getattr(_layers, k) for k in dir(_layers) if not k.startswith("__")})
sys.modules["timm.models.layers"] = pkg


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


OUT = os.path.join(ROOT, "tests", "golden")


def _randomize_bn(m, g):
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(0.3 * torch.randn(mod.num_features, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
            mod.weight.data.copy_(0.7 + 0.6 * torch.rand(mod.num_features, generator=g))
            mod.bias.data.copy_(0.2 * torch.randn(mod.num_features, generator=g))


def main():
    mlla = _load("/root/reference/mlla/models/mlla_recconv.py", "ref_mlla_recconv")
    # name, B, C, H, W, level, mode, bias
    for name, B, C, H, W, level, mode, bias in [("mlla_l2_nearest", 2, 6, 16, 24, 2, "nearest", False), ("mlla_l3_nearest_bias", 1, 4, 32, 32, 3, "nearest", True),
                                                ("mlla_l1_bilinear", 2, 3, 10, 6, 1, "bilinear", False)]:
        torch.manual_seed(0)
        m = mlla.RecConv2d(C, kernel_size=5, bias=bias, level=level, mode=mode).eval()
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, C, H, W, generator=g)
        with torch.no_grad():
            y = m(x)
        d = dict(x=x.numpy(), y=y.numpy(), meta=np.array([B, C, H, W, level, 0 if mode == "bilinear" else 1, int(bias)]), torch_version=np.array(torch.__version__))
        for k, v in m.state_dict().items():
            d["sd:" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, f"variant_{name}.npz"), **d)
        print(name, tuple(y.shape), float(y.abs().mean()))

    ls = _load("/root/reference/lsnet/model/recattn.py", "ref_lsnet_recattn")
    # name, B, C (RecAttn2d runs on C / 4), heads, H, W, stage
    for name, B, C, heads, H, W, stage in [("lsnet_partial_s0", 1, 32, 2, 28, 28, 0), ("lsnet_partial_s1", 2, 64, 2, 14, 14, 1), ("lsnet_partial_s2_odd", 1, 64, 4, 9, 13, 2)]:
        torch.manual_seed(0)
        m = ls.PartialChannelOperation(C, ls.RecAttn2d(C // 4, num_heads=heads, stage=stage), split_rate=4)
        g = torch.Generator().manual_seed(1)
        _randomize_bn(m, g)
        m.eval()
        x = torch.randn(B, C, H, W, generator=g)
        with torch.no_grad():
            y = m(x)
        d = dict(x=x.numpy(), y=y.numpy(), meta=np.array([B, C, heads, H, W, stage]), torch_version=np.array(torch.__version__))
        for k, v in m.state_dict().items():
            d["sd:" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, f"variant_{name}.npz"), **d)
        print(name, tuple(y.shape), float(y.abs().mean()))


if __name__ == "__main__":
    main()
