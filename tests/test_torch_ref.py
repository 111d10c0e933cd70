"""Pins oracle/torch_ref.py (the PyTorch-eager restatement used as CPU baseline) to the reference fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle.torch_ref import RefRecConv2d, recconv_reference
from tests.helpers import GOLDEN, TOL_FP32, load_recconv_golden, recconv_golden_files, rel_err


@pytest.mark.parametrize("path", recconv_golden_files(), ids=lambda p: os.path.basename(p)[8:-4])
def test_torch_ref_matches_reference_fixture(path):
    g = load_recconv_golden(path)
    z, p = g["z"], g["params"]
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    x = t(z["x"]).requires_grad_(True)
    ws = [t(w).requires_grad_(True) for w in [p.down_w, *p.convs_w]]
    bs = [t(b).requires_grad_(True) for b in [p.down_b, *p.convs_b]] if g["bias"] else None
    y = recconv_reference(x, ws[0], ws[1:], bs[0] if bs else None, bs[1:] if bs else None, g["mode"])
    assert rel_err(y.detach().numpy(), z["y"]) < TOL_FP32
    y.backward(t(z["gy"]))
    assert rel_err(x.grad.numpy(), z["gx"]) < TOL_FP32
    for j in range(g["L"] + 1):
        assert rel_err(ws[1 + j].grad.numpy(), z[f"g:convs.{j}.weight"]) < TOL_FP32


def test_ref_module_state_dict_layout():
    m = RefRecConv2d(8, kernel_size=5, bias=True, level=3)
    keys = list(m.state_dict().keys())
    assert keys == ["down.weight", "down.bias"] + [f"convs.{j}.{n}" for j in range(4) for n in ("weight", "bias")]
    assert tuple(m.down.weight.shape) == (8, 1, 5, 5)


def test_ref_recattn_matches_reference_fixtures():
    """oracle.torch_ref.RefRecAttn2d (the CPU baseline of bench.py --model recnext_a3) against outputs of the
    unmodified reference RecAttn2d (tests/golden/recattn_*.npz, oracle/gen_golden_recattn.py)."""
    import glob

    from oracle.torch_ref import RefRecAttn2d

    files = sorted(glob.glob(os.path.join(GOLDEN, "recattn_*.npz")))
    assert files
    for p in files:
        z = np.load(p)
        B, dim, heads, H, W, stage, mode = (int(v) for v in z["meta"])
        m = RefRecAttn2d(dim, heads, stage=stage, mode=["bilinear", "nearest"][mode])
        m.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd:")}, strict=True)
        with torch.no_grad():
            y = m.eval()(torch.from_numpy(z["x"])).numpy()
        assert rel_err(y, z["y"]) < 1e-6
