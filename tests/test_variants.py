"""The reference's other wirings of the hot path: the MLLA ablation RecConv2d (mlla/models/mlla_recconv.py:20-50) and the
L-series partial-channel RecAttn2d (lsnet/model/recattn.py:115-127, :226-237), against fixtures produced by the UNMODIFIED reference
modules (oracle/gen_golden_variants.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MLLA = sorted(glob.glob(os.path.join(GOLD, "variant_mlla_*.npz")))
LSNET = sorted(glob.glob(os.path.join(GOLD, "variant_lsnet_*.npz")))


def _sd(d):
    return {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd:")}


def test_variant_fixtures_exist():
    assert len(MLLA) == 3 and len(LSNET) == 3


@pytest.mark.parametrize("path", MLLA, ids=lambda p: os.path.basename(p)[8:-4])
def test_mlla_recconv_equals_size_interpolate_cpu(path):
    """host logic + oracle: with every level halving exactly, the scale-factor-2 up path IS interpolate(size=...), so the restatement of the
    main RecConv2d (oracle/torch_ref.py) reproduces the unmodified MLLA module; and the state_dict loads strictly"""
    from oracle.torch_ref import RefRecConv2d
    from recnext_b200.variants import MllaRecConv2d

    d = np.load(path)
    B, C, H, W, level, mode, bias = (int(v) for v in d["meta"])
    mode = "bilinear" if mode == 0 else "nearest"
    ref = RefRecConv2d(C, kernel_size=5, bias=bool(bias), level=level, mode=mode)
    ref.load_state_dict(_sd(d), strict=True)
    with torch.no_grad():
        y = ref(torch.from_numpy(d["x"]))
    assert rel_err(y.numpy(), d["y"]) < 1e-6
    m = MllaRecConv2d(C, kernel_size=5, bias=bool(bias), level=level, mode=mode)
    m.load_state_dict(_sd(d), strict=True)
    assert sorted(m.state_dict().keys()) == sorted(_sd(d).keys())
    with pytest.raises(ValueError, match="divisible"):
        m(torch.zeros(1, C, H + 1, W))


@pytest.mark.gpu
@pytest.mark.parametrize("path", MLLA, ids=lambda p: os.path.basename(p)[8:-4])
def test_mlla_recconv_kernel_matches_reference(path):
    from recnext_b200.variants import MllaRecConv2d

    d = np.load(path)
    B, C, H, W, level, mode, bias = (int(v) for v in d["meta"])
    m = MllaRecConv2d(C, kernel_size=5, bias=bool(bias), level=level, mode="bilinear" if mode == 0 else "nearest").cuda()
    m.load_state_dict(_sd(d), strict=True)
    x = torch.from_numpy(d["x"]).cuda().requires_grad_(True)
    y = m(x)
    assert rel_err(y.detach().cpu().numpy(), d["y"]) < 1e-5          # fp32 bar (north_star)
    y.square().sum().backward()                                      # the module trains like RecConv2d: gradients flow
    assert x.grad is not None and all(p.grad is not None for p in m.parameters())
    with torch.no_grad():
        yb = m(x.detach().bfloat16()).float()
    assert rel_err(yb.cpu().numpy(), d["y"]) < 2e-2


@pytest.mark.parametrize("path", LSNET, ids=lambda p: os.path.basename(p)[8:-4])
def test_lsnet_partial_state_dict_cpu(path):
    from recnext_b200.variants import LsRecAttn2d, PartialChannelOperation

    d = np.load(path)
    B, C, heads, H, W, stage = (int(v) for v in d["meta"])
    m = PartialChannelOperation(C, LsRecAttn2d(C // 4, num_heads=heads, stage=stage), split_rate=4)
    m.load_state_dict(_sd(d), strict=True)                           # same module tree and keys as the reference
    m.eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(torch.from_numpy(d["x"]))                                  # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("path", LSNET, ids=lambda p: os.path.basename(p)[8:-4])
def test_lsnet_partial_kernel_matches_reference(path):
    from recnext_b200.variants import LsRecAttn2d, PartialChannelOperation

    d = np.load(path)
    B, C, heads, H, W, stage = (int(v) for v in d["meta"])
    m = PartialChannelOperation(C, LsRecAttn2d(C // 4, num_heads=heads, stage=stage), split_rate=4).cuda().eval()
    m.load_state_dict(_sd(d), strict=True)
    x = torch.from_numpy(d["x"]).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):   # the RecAttn2d kernels are 16-bit (inference under autocast)
        y = m(x).float().cpu().numpy()
    ref = d["y"]
    q = C // 4
    assert rel_err(y[:, :q], ref[:, :q]) < 3e-2                      # the RecAttn2d quarter: 16-bit kernels against the fp32 reference
    assert np.array_equal(y[:, q:], x.bfloat16().float().cpu().numpy()[:, q:])   # the other channels pass through untouched
