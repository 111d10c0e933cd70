"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference fixtures.

Bars (BASELINE.json north_star): relative error <= 1e-5 in fp32, <= 2e-2 in bf16 (max-norm relative, see
tests/helpers.rel_err); interpolation source indices bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import recconv_oracle as O
from oracle.torch_ref import RefRecConv2d, recconv_reference
from tests.helpers import GOLDEN, TOL_BF16, TOL_FP32, load_recconv_golden, recconv_golden_files, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _R():
    import recnext_b200 as R

    return R


def _lists(p, dev=DEV, dtype=torch.float32):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype)  # noqa: E731
    ws = [t(p.down_w)] + [t(w) for w in p.convs_w]
    bs = ([t(p.down_b)] + [t(b) for b in p.convs_b]) if p.down_b is not None else None
    return ws, bs


def _run(x, gy, p, mode, dtype):
    R = _R()
    ws, bs = _lists(p)
    xd = torch.from_numpy(x).to(DEV, dtype)
    y = R.recconv_forward(xd, ws, bs, p.k, p.level, mode)
    gx, gw, gb = R.recconv_backward(xd, torch.from_numpy(gy).to(DEV, dtype), ws, bs, p.k, p.level, mode)
    torch.cuda.synchronize()
    return y.float().cpu().numpy(), gx.float().cpu().numpy(), gw.cpu().numpy(), None if gb is None else gb.cpu().numpy()


def _compare(y, gx, gw, gb, ref_y, ref, p, tol):
    C, k, L = p.down_w.shape[0], p.k, p.level
    assert rel_err(y, ref_y) < tol
    assert rel_err(gx, ref["gx"]) < tol
    if L > 0:
        assert rel_err(gw[0].reshape(C, 1, k, k), ref["down_w"]) < tol
    for j in range(L + 1):
        assert rel_err(gw[1 + j].reshape(C, 1, k, k), ref["convs_w"][j]) < tol
    if gb is not None:
        if L > 0:
            assert rel_err(gb[0], ref["down_b"]) < tol
        for j in range(L + 1):
            assert rel_err(gb[1 + j], ref["convs_b"][j]) < tol


_FIX = [p for p in recconv_golden_files() if "100x167" not in p]


@pytest.mark.parametrize("path", _FIX, ids=lambda p: os.path.basename(p)[8:-4])
def test_fixture_fp32(path):
    """Outputs and all gradients of the UNMODIFIED reference (tests/golden, torch CPU fp32)."""
    g = load_recconv_golden(path)
    z, p = g["z"], g["params"]
    y, gx, gw, gb = _run(z["x"], z["gy"], p, g["mode"], torch.float32)
    ref = dict(gx=z["gx"], down_w=z["g:down.weight"], convs_w=[z[f"g:convs.{j}.weight"] for j in range(g["L"] + 1)])
    if g["bias"]:
        ref["down_b"] = z["g:down.bias"]
        ref["convs_b"] = [z[f"g:convs.{j}.bias"] for j in range(g["L"] + 1)]
    _compare(y, gx, gw, gb, z["y"], ref, p, TOL_FP32)


@pytest.mark.parametrize("path", _FIX, ids=lambda p: os.path.basename(p)[8:-4])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_fixture_low_precision(path, dtype):
    g = load_recconv_golden(path)
    z, p = g["z"], g["params"]
    y, gx, gw, gb = _run(z["x"], z["gy"], p, g["mode"], dtype)
    tol = TOL_BF16 if dtype == torch.bfloat16 else 4e-3
    ref = dict(gx=z["gx"], down_w=z["g:down.weight"], convs_w=[z[f"g:convs.{j}.weight"] for j in range(g["L"] + 1)])
    if g["bias"]:
        ref["down_b"] = z["g:down.bias"]
        ref["convs_b"] = [z[f"g:convs.{j}.bias"] for j in range(g["L"] + 1)]
    _compare(y, gx, gw, gb, z["y"], ref, p, tol)
    # and the reference's own bf16 autocast output (every intermediate rounded) is within the same bar
    if dtype == torch.bfloat16:
        assert rel_err(y, z["y_bf16_autocast"]) < TOL_BF16


def test_384px_stage0_forward():
    """96x96 planes (384-px input, stage 0): forward fits on chip; the fp32 backward pyramid does not (yet)."""
    rng = np.random.default_rng(5)
    p = O.RecConvParams.random(16, 5, 4, False, rng)
    x = rng.standard_normal((1, 16, 96, 96), dtype=np.float32)
    ws, bs = _lists(p)
    y = _R().recconv_forward(torch.from_numpy(x).to(DEV), ws, bs, 5, 4, "bilinear")
    assert rel_err(y.cpu().numpy(), O.forward(x, p, "bilinear")) < TOL_FP32


def test_fixture_big_plane_forward_backward():
    """Detection stage 1, unpadded odd width (100x167, level 3): outputs and every gradient of the unmodified reference.
    The fp32 backward pyramid does not fit on chip: the streamed path (csrc/gstream.cu) takes it."""
    g = load_recconv_golden(os.path.join(GOLDEN, "recconv_det_odd_100x167_L3.npz"))
    z, p = g["z"], g["params"]
    y, gx, gw, gb = _run(z["x"], z["gy"], p, g["mode"], torch.float32)
    ref = dict(gx=z["gx"], down_w=z["g:down.weight"], convs_w=[z[f"g:convs.{j}.weight"] for j in range(g["L"] + 1)])
    _compare(y, gx, gw, gb, z["y"], ref, p, TOL_FP32)
    y, gx, gw, gb = _run(z["x"], z["gy"], p, g["mode"], torch.bfloat16)
    _compare(y, gx, gw, gb, z["y"], ref, p, TOL_BF16)


@pytest.fixture
def streamed_path():
    """Forces the level-by-level streamed kernels (the path of planes whose pyramid does not fit on chip)."""
    old = os.environ.get("RECNEXT_PATH")
    os.environ["RECNEXT_PATH"] = "stream"
    yield
    if old is None:
        del os.environ["RECNEXT_PATH"]
    else:
        os.environ["RECNEXT_PATH"] = old


@pytest.mark.parametrize("path", recconv_golden_files(), ids=lambda p: os.path.basename(p)[8:-4])
def test_streamed_path_fixture_fp32(path, streamed_path):
    """Every reference fixture (k = 3/5/7, nearest, bias, odd sizes, level 0..4) through the streamed kernels."""
    g = load_recconv_golden(path)
    z, p = g["z"], g["params"]
    assert "streamed" in _R().plan_describe(z["x"].shape, p.k, p.level, g["mode"], torch.float32, g["bias"], True)
    y, gx, gw, gb = _run(z["x"], z["gy"], p, g["mode"], torch.float32)
    ref = dict(gx=z["gx"], down_w=z["g:down.weight"], convs_w=[z[f"g:convs.{j}.weight"] for j in range(g["L"] + 1)])
    if g["bias"]:
        ref["down_b"] = z["g:down.bias"]
        ref["convs_b"] = [z[f"g:convs.{j}.bias"] for j in range(g["L"] + 1)]
    _compare(y, gx, gw, gb, z["y"], ref, p, TOL_FP32)


# BASELINE configs[4]: RecNeXt-M3 backbone at detection scale 800x1333 (padded to 800x1344), 2 images per GPU
_DET_CASES = [
    (2, 64, 200, 336, 4),    # stage 0
    (2, 128, 100, 168, 3),   # stage 1
    (2, 64, 200, 334, 4),    # stage 0, unpadded 1333-px width (odd sizes down the pyramid: 334 -> 167 -> 84 -> 42 -> 21)
]


@pytest.mark.parametrize("case", _DET_CASES, ids=lambda c: "x".join(str(v) for v in c))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
def test_detection_scale_forward_backward(case, dtype):
    """Forward and ALL gradients at the detection plane sizes against the C oracle (fp64 tap sums)."""
    B, C, H, W, L = case
    rng = np.random.default_rng(7)
    p = O.RecConvParams.random(C, 5, L, False, rng)
    x = rng.standard_normal((B, C, H, W), dtype=np.float32)
    gy = rng.standard_normal((B, C, H, W), dtype=np.float32)
    if dtype == torch.bfloat16:
        x = torch.from_numpy(x).bfloat16().float().numpy()
        gy = torch.from_numpy(gy).bfloat16().float().numpy()
    y, gx, gw, gb = _run(x, gy, p, "bilinear", dtype)
    tol = TOL_BF16 if dtype == torch.bfloat16 else TOL_FP32
    _compare(y, gx, gw, gb, O.forward(x, p, "bilinear"), O.backward(x, gy, p, "bilinear"), p, tol)


# BASELINE.json per-stage shapes at reduced batch, plus ragged / odd / tiny cases.
_ORACLE_CASES = [
    # B, C, H, W, L, k, mode, bias
    (3, 64, 56, 56, 4, 5, "bilinear", False),     # M3 stage 0
    (3, 128, 28, 28, 3, 5, "bilinear", False),    # M3 stage 1
    (5, 256, 14, 14, 2, 5, "bilinear", False),    # M3 stage 2
    (5, 512, 7, 7, 1, 5, "bilinear", False),      # M3 stage 3
    (2, 40, 56, 56, 4, 5, "bilinear", True),      # M0 stage 0 (+bias: fused-BN eval form)
    (2, 80, 56, 56, 4, 5, "nearest", False),      # M5 width, nearest
    (2, 256, 50, 84, 2, 5, "bilinear", False),    # detection stage 2
    (2, 512, 25, 42, 1, 5, "bilinear", True),     # detection stage 3
    (1, 7, 25, 21, 3, 5, "bilinear", True),       # odd channels: ragged channel groups, no TMA
    (2, 3, 13, 9, 2, 3, "nearest", True),
    (2, 5, 31, 17, 2, 7, "bilinear", False),
    (1, 1, 1, 1, 2, 5, "bilinear", True),         # degenerate 1x1 plane
    (4, 6, 2, 3, 1, 5, "nearest", False),
    (7, 9, 8, 8, 0, 5, "bilinear", True),         # level 0: plain depthwise conv
]


@pytest.mark.parametrize("case", _ORACLE_CASES, ids=lambda c: "x".join(str(v) for v in c))
def test_against_oracle_fp32(case):
    B, C, H, W, L, k, mode, bias = case
    rng = np.random.default_rng(hash(case) & 0xFFFF)
    p = O.RecConvParams.random(C, k, L, bias, rng)
    x = rng.standard_normal((B, C, H, W), dtype=np.float32)
    gy = rng.standard_normal((B, C, H, W), dtype=np.float32)
    y, gx, gw, gb = _run(x, gy, p, mode, torch.float32)
    _compare(y, gx, gw, gb, O.forward(x, p, mode), O.backward(x, gy, p, mode), p, TOL_FP32)


@pytest.mark.parametrize("case", _ORACLE_CASES[:8], ids=lambda c: "x".join(str(v) for v in c))
def test_against_oracle_bf16(case):
    B, C, H, W, L, k, mode, bias = case
    rng = np.random.default_rng(hash(case) & 0xFFFF)
    p = O.RecConvParams.random(C, k, L, bias, rng)
    x = torch.from_numpy(rng.standard_normal((B, C, H, W), dtype=np.float32)).bfloat16().float().numpy()
    gy = torch.from_numpy(rng.standard_normal((B, C, H, W), dtype=np.float32)).bfloat16().float().numpy()
    y, gx, gw, gb = _run(x, gy, p, mode, torch.bfloat16)
    _compare(y, gx, gw, gb, O.forward(x, p, mode), O.backward(x, gy, p, mode), p, TOL_BF16)


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
@pytest.mark.parametrize("hw", [(7, 7), (14, 14), (25, 42), (13, 21), (167, 84), (9, 257)])
def test_interpolation_source_indices_bit_exact(mode, hw):
    """Identity filters turn the block into y = x + up(x[::2, ::2]); with an index-valued x the result exposes
    the kernel's source indices, compared bit-for-bit with the oracle (whose tables are pinned to ATen)."""
    H, W = hw
    k, L, C = 5, 1, 2
    delta = np.zeros((C, 1, k, k), np.float32)
    delta[:, 0, k // 2, k // 2] = 1.0
    p = O.RecConvParams(down_w=delta, convs_w=[delta.copy(), delta.copy()])
    Hl, Wl = (H + 1) // 2, (W + 1) // 2
    x = np.zeros((1, C, H, W), np.float32)
    rows = np.arange(Hl, dtype=np.float32)[:, None] * np.ones((1, Wl), np.float32)
    cols = np.ones((Hl, 1), np.float32) * np.arange(Wl, dtype=np.float32)[None, :]
    x[0, 0, ::2, ::2] = rows   # channel 0 exposes the row source coordinate
    x[0, 1, ::2, ::2] = cols   # channel 1 exposes the column source coordinate
    R = _R()
    ws, bs = _lists(p)
    y = R.recconv_forward(torch.from_numpy(x).to(DEV), ws, bs, k, L, mode).cpu().numpy()
    up = y - x  # exact: x is integer-valued and |up| < 2^10
    for ch, (n_in, n_out, axis) in enumerate([(Hl, H, 0), (Wl, W, 1)]):
        line = up[0, ch].take(0, axis=1 - axis)  # first column (rows) / first row (cols)
        for d in range(n_out):
            if mode == "nearest":
                assert line[d] == float(O.nearest_index(n_in, n_out, d)), (ch, d)
            else:
                i0, i1, lam = O.bilinear_index(n_in, n_out, d)
                expect = np.float32(np.float32(1.0 - lam) * np.float32(i0) + np.float32(lam) * np.float32(i1))
                assert abs(line[d] - expect) <= 2.0 ** -21 * max(1.0, expect), (ch, d, line[d], expect)
                if 1e-4 < lam < 1 - 1e-4:
                    assert int(np.floor(line[d])) == i0, (ch, d)


def test_module_autograd_matches_reference_module():
    """nn.Module drop-in: same state_dict, forward/backward through autograd vs the PyTorch restatement on GPU."""
    R = _R()
    torch.manual_seed(0)
    m = R.RecConv2d(32, kernel_size=5, bias=True, level=3).to(DEV)
    ref = RefRecConv2d(32, kernel_size=5, bias=True, level=3).to(DEV)
    ref.load_state_dict(m.state_dict(), strict=True)
    x = torch.randn(4, 32, 28, 28, device=DEV, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y, yr = m(x), ref(xr)
        gy = torch.randn_like(y)
        y.backward(gy); yr.backward(gy)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert rel_err(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < TOL_FP32
    assert rel_err(x.grad.cpu().numpy(), xr.grad.cpu().numpy()) < TOL_FP32
    for (n, a), (_, b) in zip(m.named_parameters(), ref.named_parameters()):
        assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 2e-5, n


def test_module_autocast_bf16():
    R = _R()
    torch.manual_seed(1)
    m = R.RecConv2d(64, level=4).to(DEV)
    ref = RefRecConv2d(64, level=4).to(DEV)
    ref.load_state_dict(m.state_dict())
    x = torch.randn(2, 64, 56, 56, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    assert y.dtype == torch.bfloat16
    yr = ref(x)  # fp32 eager
    assert rel_err(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy()) < TOL_BF16


@pytest.mark.parametrize("shape,level", [((256, 64, 56, 56), 4), ((256, 256, 14, 14), 2), ((256, 512, 7, 7), 1)],
                         ids=["stage0", "stage2", "stage3"])
def test_full_size_properties_bf16(shape, level):
    """BASELINE config-2 sizes (M3, batch 256, bf16): compare with the fp32 PyTorch restatement on the GPU and
    check size-independent properties: batch-slice consistency and determinism of the gradient reduction."""
    R = _R()
    B, C, H, W = shape
    torch.manual_seed(2)
    m = R.RecConv2d(C, level=level).to(DEV)
    x = torch.randn(shape, device=DEV).bfloat16()
    gy = torch.randn(shape, device=DEV).bfloat16()
    ws, bs = m._param_lists()
    ws = [w.detach() for w in ws]
    y = R.recconv_forward(x, ws, None, 5, level, "bilinear")
    gx, gw, _ = R.recconv_backward(x, gy, ws, None, 5, level, "bilinear")
    # (1) every image is independent: a batch slice gives bit-identical rows
    sl = slice(100, 117)
    y_s = R.recconv_forward(x[sl].contiguous(), ws, None, 5, level, "bilinear")
    assert torch.equal(y[sl], y_s)
    gx_s, gw_s, _ = R.recconv_backward(x[sl].contiguous(), gy[sl].contiguous(), ws, None, 5, level, "bilinear")
    assert torch.equal(gx[sl], gx_s)
    # (2) deterministic
    gx2, gw2, _ = R.recconv_backward(x, gy, ws, None, 5, level, "bilinear")
    assert torch.equal(gx, gx2) and torch.equal(gw, gw2)
    # (3) against the fp32 PyTorch restatement on the same inputs (subset of the batch to bound memory/time)
    nb = 32
    xr = x[:nb].float().requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        yr = recconv_reference(xr, wr[0], wr[1:], None, None, "bilinear")
        yr.backward(gy[:nb].float())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert rel_err(y[:nb].float().cpu().numpy(), yr.detach().cpu().numpy()) < TOL_BF16
    assert rel_err(gx[:nb].float().cpu().numpy(), xr.grad.cpu().numpy()) < TOL_BF16
    _, gw_nb, _ = R.recconv_backward(x[:nb].contiguous(), gy[:nb].contiguous(), ws, None, 5, level, "bilinear")
    for j in range(level + 2):
        assert rel_err(gw_nb[j].view(C, 1, 5, 5).cpu().numpy(), wr[j].grad.cpu().numpy()) < TOL_BF16, j
    # (4) linearity in x (fp32 path, same weights): f(a*x1 + x2) == a*f(x1) + f(x2)
    x1 = torch.randn(8, C, H, W, device=DEV); x2 = torch.randn(8, C, H, W, device=DEV)
    f = lambda t: R.recconv_forward(t, ws, None, 5, level, "bilinear")  # noqa: E731
    lhs, rhs = f(0.5 * x1 + x2), 0.5 * f(x1) + f(x2)
    assert rel_err(lhs.cpu().numpy(), rhs.cpu().numpy()) < TOL_FP32


def test_empty_batch_and_noncontiguous_input():
    R = _R()
    m = R.RecConv2d(8, level=2).to(DEV)
    assert m(torch.empty(0, 8, 14, 14, device=DEV)).shape == (0, 8, 14, 14)
    x = torch.randn(2, 14, 14, 8, device=DEV).permute(0, 3, 1, 2)  # channels_last-style view
    ref = RefRecConv2d(8, level=2).to(DEV)
    ref.load_state_dict(m.state_dict())
    assert rel_err(m(x).detach().cpu().numpy(), ref(x).detach().cpu().numpy()) < TOL_FP32


def test_runs_on_side_stream_and_under_cuda_graph():
    R = _R()
    m = R.RecConv2d(16, level=2).to(DEV)
    x = torch.randn(4, 16, 14, 14, device=DEV)
    ws = [w.detach() for w in m._param_lists()[0]]
    y0 = R.recconv_forward(x, ws, None, 5, 2, "bilinear")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y1 = R.recconv_forward(x, ws, None, 5, 2, "bilinear")
    s.synchronize()
    assert torch.equal(y0, y1)
    g = torch.cuda.CUDAGraph()
    ybuf = None
    with torch.cuda.graph(g):
        ybuf = R.recconv_forward(x, ws, None, 5, 2, "bilinear")
    x.copy_(torch.randn_like(x))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(ybuf, R.recconv_forward(x, ws, None, 5, 2, "bilinear"))
