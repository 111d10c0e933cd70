"""Generates tests/golden/recattn_*.npz by running the UNMODIFIED reference RecAttn2d (model/recattn.py:54-67) in this
container (through oracle/timm_shim; timm is third-party and absent).  TEST INFRASTRUCTURE.

    python oracle/gen_golden_recattn.py

For each small case: eval-mode module with non-trivial BatchNorm statistics, input, fp32 output, the intermediate
results of the two pieces the CUDA kernels replace (the fused stride-2 ConvNorm and the `conv(x + interpolate(z))`
tail) and the full state_dict — which also pins the state_dict layout of recnext_b200.RecAttn2d.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
sys.path.insert(0, "/root/reference")

from model.recattn import RecAttn2d  # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "tests", "golden")
# name, B, dim, heads, H, W, stage, mode
CASES = [("a_stage0_56", 1, 16, 2, 56, 56, 0, "nearest"), ("a_stage1_28", 2, 32, 4, 28, 28, 1, "nearest"),
         ("a_stage2_14", 3, 32, 8, 14, 14, 2, "nearest"), ("a_stage3_7", 3, 64, 16, 7, 7, 3, "nearest"),
         ("a_odd_25x21_bilinear", 1, 8, 2, 25, 21, 1, "bilinear")]


def main():
    for name, B, dim, heads, H, W, stage, mode in CASES:
        torch.manual_seed(0)
        m = RecAttn2d(dim, heads, stage=stage, mode=mode)
        g = torch.Generator().manual_seed(1)
        for mod in m.modules():  # non-trivial BatchNorm statistics and affine parameters
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(0.3 * torch.randn(mod.num_features, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
                mod.weight.data.copy_(0.7 + 0.6 * torch.rand(mod.num_features, generator=g))
                mod.bias.data.copy_(0.2 * torch.randn(mod.num_features, generator=g))
        m.eval()
        x = torch.randn(B, dim, H, W, generator=g)
        with torch.no_grad():
            low = m.down[0](x)
            z = m.down[1](low)
            y = m(x)
        d = dict(x=x.numpy(), low=low.numpy(), z=z.numpy(), y=y.numpy(),
                 meta=np.array([B, dim, heads, H, W, stage, 0 if mode == "bilinear" else 1]), torch_version=np.array(torch.__version__))
        for k, v in m.state_dict().items():
            d["sd:" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, f"recattn_{name}.npz"), **d)
        print(name, tuple(y.shape), float(y.abs().mean()))




def model_logits_a():
    """RecNeXt-A0 eval logits with the name-keyed deterministic weights of oracle/detinit.py (reference model/recattn.py)."""
    sys.path.insert(0, ROOT)
    from timm.models import create_model  # shim

    from oracle.detinit import fill_state_dict_

    d = {}
    net = create_model("recnext_a0").eval()
    fill_state_dict_(net, seed=0)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 224, 224, generator=g)
    with torch.no_grad():
        d["recnext_a0_224_logits"] = net(x).numpy()
        d["recnext_a0_224_feat_mean"] = net.forward_features(x).mean((2, 3)).numpy()
    d["recnext_a0_224_nparams"] = np.array(sum(p.numel() for p in net.parameters()))
    d["recnext_a0_keys"] = np.array(sorted(net.state_dict().keys()))
    d["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(OUT, "model_logits_a.npz"), **d)
    print("recnext_a0 logits", d["recnext_a0_224_logits"][0, :4])


if __name__ == "__main__":
    main()
    model_logits_a()
