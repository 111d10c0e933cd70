"""Generates tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE.  Run once here (the reference does not exist on the GPU box):

    python oracle/gen_golden.py

Imports /root/reference/model/recnext.py through oracle/timm_shim (timm is third-party and not
installed) and records, for a set of small cases, the inputs, weights, outputs and all gradients
of reference `RecConv2d` (model/recnext.py:8-34), F.interpolate source-index tables, and RecNeXt-M0
logits under the deterministic name-keyed init of oracle/detinit.py.  torch version is stored in
every file.  The committed fixtures are what pins oracle/recconv_oracle.c and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from model.recnext import RecConv2d  # noqa: E402  (the reference)
from timm.models import create_model  # noqa: E402  (shim)

from oracle.detinit import fill_state_dict_  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, B, C, H, W, level, k, mode, bias)
RECCONV_CASES = [
    ("m_stage0_56_L4", 1, 3, 56, 56, 4, 5, "bilinear", False),
    ("m_stage1_28_L3", 2, 3, 28, 28, 3, 5, "bilinear", False),
    ("m_stage2_14_L2_nearest", 2, 4, 14, 14, 2, 5, "nearest", False),
    ("m_stage3_7_L1_bias", 2, 5, 7, 7, 1, 5, "bilinear", True),
    ("det_stage3_25x42_L1_bias", 1, 3, 25, 42, 1, 5, "bilinear", True),
    ("det_stage2_50x84_L2", 1, 2, 50, 84, 2, 5, "bilinear", False),
    ("det_odd_100x167_L3", 1, 1, 100, 167, 3, 5, "bilinear", False),
    ("k3_33x17_L3_nearest_bias", 2, 2, 33, 17, 3, 3, "nearest", True),
    ("k7_40x31_L2", 1, 2, 40, 31, 2, 7, "bilinear", False),
    ("tiny_5x3_L4_bias", 2, 3, 5, 3, 4, 5, "bilinear", True),
    ("level0_9x11", 1, 2, 9, 11, 0, 5, "bilinear", True),
    ("nearest_odd_23x29_L2", 1, 2, 23, 29, 2, 5, "nearest", True),
]


def recconv_case(name, B, C, H, W, L, k, mode, bias, seed=0):
    torch.manual_seed(seed)
    m = RecConv2d(C, kernel_size=k, bias=bias, level=L, mode=mode)
    x = torch.randn(B, C, H, W, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    d = dict(x=x.detach().numpy(), y=y.detach().numpy(), gy=gy.numpy(), gx=x.grad.numpy())
    for n, p in m.named_parameters():
        d["w:" + n] = p.detach().numpy()
        d["g:" + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()  # level 0: `down` unused
    # the same module under CPU bf16 autocast (every intermediate rounded to bf16)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        d["y_bf16_autocast"] = m(x.detach().bfloat16()).float().numpy()
    d["meta"] = np.array([B, C, H, W, L, k, {"bilinear": 0, "nearest": 1}[mode], int(bias)], np.int64)
    d["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(OUT, f"recconv_{name}.npz"), **d)
    return float(y.detach().sum()), float(y.detach().abs().sum())


def index_tables():
    """F.interpolate on index ramps: the values ARE the source coordinates (bit-exact check)."""
    pairs = [(4, 7), (7, 14), (14, 28), (28, 56), (13, 25), (21, 42), (25, 50), (42, 84), (84, 167), (84, 168),
             (3, 5), (2, 3), (1, 2), (1, 3), (1, 1), (5, 9), (6, 11), (12, 23), (15, 29), (8, 15), (100, 200),
             (129, 257)]  # 129->257: the one pair below 400 where fused vs unfused src differ (dst 128)
    d = {}
    for i, o in pairs:
        ramp = torch.arange(i, dtype=torch.float32).view(1, 1, 1, i)
        d[f"bilinear_{i}_{o}"] = torch.nn.functional.interpolate(ramp, size=(1, o), mode="bilinear").view(-1).numpy()
        d[f"nearest_{i}_{o}"] = torch.nn.functional.interpolate(ramp, size=(1, o), mode="nearest").view(-1).numpy().astype(np.int32)
    d["pairs"] = np.array(pairs, np.int32)
    d["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(OUT, "interp_index_tables.npz"), **d)


def survey_fingerprints():
    """SURVEY.md §8(c) known-answer sums, regenerated (construct module THEN x; gy = 0.5)."""
    rows = []
    for C, H, W, L, mode, dt in [(8, 56, 56, 4, "bilinear", torch.float32), (8, 25, 42, 1, "bilinear", torch.float32),
                                 (8, 14, 14, 2, "nearest", torch.float32)]:
        torch.manual_seed(0)
        m = RecConv2d(C, 5, level=L, mode=mode)
        x = torch.randn(2, C, H, W, requires_grad=True)
        y = m(x)
        y.backward(0.5 * torch.ones_like(y))
        rows.append([C, H, W, L, {"bilinear": 0, "nearest": 1}[mode], float(y.sum()), float(y.abs().sum()),
                     float(x.grad.abs().sum()), float(m.down.weight.grad.abs().sum()),
                     float(m.convs[0].weight.grad.abs().sum())])
    np.savez_compressed(os.path.join(OUT, "survey_fingerprints.npz"), rows=np.array(rows, np.float64),
                        torch_version=np.array(torch.__version__))
    return rows


def model_logits():
    """RecNeXt-M0 / M3 eval logits with name-keyed deterministic weights (oracle/detinit.py)."""
    d = {}
    for variant, res in [("recnext_m0", 224), ("recnext_m3", 224), ("recnext_m0", 160)]:
        net = create_model(variant).eval()
        fill_state_dict_(net, seed=0)
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(2, 3, res, res, generator=g)
        with torch.no_grad():
            d[f"{variant}_{res}_logits"] = net(x).numpy()
            feats = net.forward_features(x)
            d[f"{variant}_{res}_feat_mean"] = feats.mean((2, 3)).numpy()
        d[f"{variant}_{res}_nparams"] = np.array(sum(p.numel() for p in net.parameters()))
        d[f"{variant}_keys"] = np.array(sorted(net.state_dict().keys()))
    d["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(OUT, "model_logits.npz"), **d)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for case in RECCONV_CASES:
        print(case[0], recconv_case(*case))
    index_tables()
    print(survey_fingerprints())
    model_logits()
    print("wrote", sorted(os.listdir(OUT)))
