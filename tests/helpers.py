"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity bars from BASELINE.json north_star: relative error 1e-5 (fp32), 2e-2 (bf16).
# "Relative" = max |a-b| / max |b| over the tensor (a per-element ratio is meaningless at zero crossings).
TOL_FP32 = 1e-5
TOL_BF16 = 2e-2


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def recconv_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "recconv_*.npz")))


def load_recconv_golden(path):
    """-> dict with meta fields + oracle RecConvParams built from the stored reference weights."""
    from oracle.recconv_oracle import RecConvParams

    z = np.load(path)
    B, C, H, W, L, k, mode, bias = (int(v) for v in z["meta"])
    p = RecConvParams(
        down_w=z["w:down.weight"],
        convs_w=[z[f"w:convs.{j}.weight"] for j in range(L + 1)],
        down_b=z["w:down.bias"] if bias else None,
        convs_b=[z[f"w:convs.{j}.bias"] for j in range(L + 1)] if bias else None,
    )
    return dict(z=z, B=B, C=C, H=H, W=W, L=L, k=k, mode=["bilinear", "nearest"][mode], bias=bool(bias), params=p)
