"""RecAttn2d (A-series token mixer, reference model/recattn.py:54-67): state_dict layout and host logic on CPU, parity of
the two CUDA pieces and of the whole module against fixtures generated from the UNMODIFIED reference
(oracle/gen_golden_recattn.py) on the GPU."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import GOLDEN, TOL_BF16, rel_err

FIX = sorted(glob.glob(os.path.join(GOLDEN, "recattn_*.npz")))


def _load(path):
    z = np.load(path)
    B, dim, heads, H, W, stage, mode = (int(v) for v in z["meta"])
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd:")}
    return z, dict(B=B, dim=dim, heads=heads, H=H, W=W, stage=stage, mode=["bilinear", "nearest"][mode]), sd


def _module(meta, sd):
    from recnext_b200.recattn import RecAttn2d

    m = RecAttn2d(meta["dim"], meta["heads"], stage=meta["stage"], mode=meta["mode"])
    m.load_state_dict(sd, strict=True)  # same keys and shapes as the reference module
    return m.eval()


def test_fixtures_exist():
    assert len(FIX) >= 5


@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[8:-4])
def test_state_dict_layout_matches_reference(path):
    _, meta, sd = _load(path)
    m = _module(meta, sd)
    assert set(m.state_dict().keys()) == set(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    # LinearAttention2 where the reference uses it (stage >= 3, model/recattn.py:58)
    assert type(m.down[1]).__name__ == ("LinearAttention2" if meta["stage"] >= 3 else "LinearAttention1")


@pytest.mark.parametrize("path", FIX[:2], ids=lambda p: os.path.basename(p)[8:-4])
def test_convnorm_fold_and_fuse_match_reference_pieces(path):
    """host logic on CPU: folding BatchNorm (model/recattn.py:87-111) reproduces the reference's ConvNorm output"""
    z, meta, sd = _load(path)
    m = _module(meta, sd)
    w, b = m.down[0].folded()
    x = torch.from_numpy(z["x"])
    low = F.conv2d(x, w, b, stride=2, padding=2, groups=meta["dim"])
    assert rel_err(low.numpy(), z["low"]) < 1e-5
    fused = m.down[0].fuse()
    assert rel_err(fused(x).detach().numpy(), z["low"]) < 1e-5
    # the product module has no CPU / eager path for the linear attention (it raises); the oracle's restatement reproduces the
    # reference's z on CPU fp32 from the same state_dict
    with pytest.raises(RuntimeError, match="CUDA"):
        m.down[1](torch.from_numpy(z["low"]))
    from oracle.torch_ref import RefLinearAttention

    ref_la = RefLinearAttention(meta["dim"], m.down[1].num_heads, quadratic=meta["stage"] >= 3).eval()
    ref_la.load_state_dict(m.down[1].state_dict(), strict=True)
    with torch.no_grad():
        zz = ref_la(torch.from_numpy(z["low"]))
    assert rel_err(zz.numpy(), z["z"]) < 1e-5


def test_no_fallbacks():
    from recnext_b200.recattn import RecAttn2d

    m = RecAttn2d(8, 2)
    with pytest.raises(RuntimeError, match="no backward"):
        m(torch.randn(1, 8, 8, 8))
    with pytest.raises(RuntimeError, match="no backward"):      # eval mode but a gradient can be asked for: no silent gradient cut
        m.eval()(torch.randn(1, 8, 8, 8))
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(torch.randn(1, 8, 8, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[8:-4])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_cuda_pieces_against_reference(path, dtype):
    from recnext_b200.recattn import recattn_down_forward, recattn_up_forward

    z, meta, sd = _load(path)
    m = _module(meta, sd).cuda()
    tol = TOL_BF16 if dtype == torch.bfloat16 else 4e-3
    x = torch.from_numpy(z["x"]).cuda().to(dtype)
    wd, bd = m.down[0].folded()
    low = recattn_down_forward(x, wd, bd)
    assert low.shape == tuple(z["low"].shape) or tuple(low.shape) == tuple(z["low"].shape)
    assert rel_err(low.float().cpu().numpy(), z["low"]) < tol
    wc, bc = m.conv.folded()
    y = recattn_up_forward(x, torch.from_numpy(z["z"]).cuda(), wc, bc, meta["mode"])
    assert rel_err(y.float().cpu().numpy(), z["y"]) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[8:-4])
def test_cuda_fp32_pieces_and_module_at_1e5(path):
    """fp32 bar of the north star (1e-5) for RecAttn2d (model/recattn.py:54-67): the two pieces on the fp32 kernels of gstream.cu and the whole
    eval-mode module (fp32 linear-attention kernel in between), against outputs of the unmodified reference"""
    from recnext_b200.recattn import recattn_down_forward, recattn_up_forward

    z, meta, sd = _load(path)
    m = _module(meta, sd).cuda()
    x = torch.from_numpy(z["x"]).cuda()
    wd, bd = m.down[0].folded()
    low = recattn_down_forward(x, wd, bd)
    assert low.dtype == torch.float32 and rel_err(low.cpu().numpy(), z["low"]) < 1e-5
    wc, bc = m.conv.folded()
    y = recattn_up_forward(x, torch.from_numpy(z["z"]).cuda(), wc, bc, meta["mode"])
    assert rel_err(y.cpu().numpy(), z["y"]) < 1e-5
    with torch.no_grad():
        ym = m(x)
    assert ym.dtype == torch.float32 and rel_err(ym.cpu().numpy(), z["y"]) < 2e-5   # (the qk / pe ConvNorms are cuDNN fp32 convs: TF32 off in conftest)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[8:-4])
def test_cuda_module_against_reference(path):
    z, meta, sd = _load(path)
    m = _module(meta, sd).cuda()
    x = torch.from_numpy(z["x"]).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    assert y.dtype == torch.bfloat16
    assert rel_err(y.float().cpu().numpy(), z["y"]) < TOL_BF16
    # fused-BN form (ConvNorm -> Conv2d with bias, what replace_batchnorm leaves): same result
    m.down[0] = m.down[0].fuse()
    m.conv = m.conv.fuse()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = m(x)
    assert rel_err(y2.float().cpu().numpy(), y.float().cpu().numpy()) < 1e-2


@pytest.mark.gpu
def test_cuda_full_size_a3_stage_shapes():
    """BASELINE config 4 (A3, batch 256 at 224^2) stage shapes: both pieces against PyTorch bf16 on the same inputs"""
    from recnext_b200.recattn import recattn_down_forward, recattn_up_forward

    for (B, C, H) in [(256, 64, 56), (256, 128, 28), (256, 256, 14), (256, 512, 7)]:
        torch.manual_seed(C)
        x = torch.randn(B, C, H, H, device="cuda").bfloat16()
        w = torch.empty(C, 1, 5, 5, device="cuda").uniform_(-0.2, 0.2)
        b = torch.empty(C, device="cuda").uniform_(-0.2, 0.2)
        low = recattn_down_forward(x, w, b)
        ref = F.conv2d(x.float(), w, b, stride=2, padding=2, groups=C)
        assert rel_err(low.float().cpu().numpy(), ref.cpu().numpy()) < TOL_BF16
        zz = torch.randn(B, C, (H + 1) // 2, (H + 1) // 2, device="cuda").bfloat16()
        y = recattn_up_forward(x, zz, w, b, "nearest")
        ref = F.conv2d(x.float() + F.interpolate(zz.float(), size=(H, H), mode="nearest"), w, b, padding=2, groups=C)
        assert rel_err(y.float().cpu().numpy(), ref.cpu().numpy()) < TOL_BF16


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[8:-4])
def test_cuda_linear_attention_kernel_against_reference(path):
    """recnext_linattn_forward (LinearAttention1/2 contraction core, model/recattn.py:21-28,44-51) against the reference's z,
    fed with the reference's own low-resolution input (fixture `low`); the two ConvNorms around it run as library convs."""
    from recnext_b200.recattn import linattn_forward

    z, meta, sd = _load(path)
    m = _module(meta, sd).cuda()
    from recnext_b200.recattn import LINATTN_HEAD_DIMS

    la = m.down[1]
    low = torch.from_numpy(z["low"]).cuda()
    assert la.head_dim in LINATTN_HEAD_DIMS
    with torch.no_grad():
        qk_pre = la.qk(low)
        pe = la.pe(low)
        out = linattn_forward(qk_pre.bfloat16(), low.bfloat16(), pe.bfloat16(), la.num_heads)
    assert rel_err(out.float().cpu().numpy(), z["z"]) < TOL_BF16
    # and through the module switch (eval, autocast): same bar
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out2 = la(low)
    assert out2.dtype == torch.bfloat16
    assert rel_err(out2.float().cpu().numpy(), z["z"]) < TOL_BF16


@pytest.mark.gpu
@pytest.mark.parametrize("d,heads,n", [(4, 3, 9), (8, 2, 70), (16, 2, 5), (20, 2, 784), (24, 4, 196), (28, 8, 49), (32, 16, 16), (40, 2, 130)])
def test_cuda_linear_attention_head_dims(d, heads, n):
    """every head_dim of the A-series (model/recattn.py:380-426) against the reference formula in fp32"""
    from recnext_b200.recattn import linattn_forward

    torch.manual_seed(d + n)
    B, dim = 3, d * heads
    qk = torch.randn(B, 2 * dim, 1, n, device="cuda").bfloat16(); v = torch.randn(B, dim, 1, n, device="cuda").bfloat16()
    pe = torch.randn(B, dim, 1, n, device="cuda").bfloat16()
    out = linattn_forward(qk, v, pe, heads)
    a = F.elu(qk.float()) + 1.0
    q, k = a.view(B, 2, heads, d, n).unbind(1)
    vv = v.float().view(B, heads, d, n)
    s = n ** -0.5
    q_t = q.transpose(-1, -2)
    kvm = (k * s) @ (vv.transpose(-1, -2) * s)
    ref = (q_t @ kvm / (q_t @ k.mean(dim=-1, keepdim=True) + 1e-6)).transpose(-1, -2).reshape(B, dim, 1, n) + pe.float()
    assert rel_err(out.float().cpu().numpy(), ref.cpu().numpy()) < TOL_BF16


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=["bf16", "f16", "f32"])
@pytest.mark.parametrize("d,heads,h,w", [(32, 2, 28, 28), (32, 4, 14, 14), (32, 8, 7, 7), (32, 16, 4, 4), (20, 2, 9, 13), (40, 2, 10, 13), (24, 3, 5, 5),
                                         (28, 2, 12, 12), (8, 2, 1, 1), (16, 2, 3, 40), (4, 3, 17, 8)])
def test_cuda_linear_attention_qk_bias_pe_entries(d, heads, h, w, dtype):
    """recnext_linattn_forward_qk / _pe (separate q / k tensors, biases added before the elu, the depthwise 3x3 `pe` conv evaluated inside
    the kernel) against the same formula in fp32 PyTorch: tensor-core kernel for 16-bit activations (every plane size class: 8-, 4- and
    1-pixel global accesses, several chunks, partial chunks, single pixels), FP32-FMA kernel for fp32"""
    from recnext_b200.recattn import linattn_forward_pe, linattn_forward_qk

    torch.manual_seed(d + h)
    B, dim, n = 3, d * heads, h * w
    q = torch.randn(B, dim, h, w, device="cuda").to(dtype); k = torch.randn_like(q); v = torch.randn_like(q)
    qb, kb = 0.3 * torch.randn(dim, device="cuda"), 0.3 * torch.randn(dim, device="cuda")
    pw, pb = 0.3 * torch.randn(dim, 1, 3, 3, device="cuda"), 0.1 * torch.randn(dim, device="cuda")
    pe = F.conv2d(v.float(), pw, pb, padding=1, groups=dim)
    r16 = (lambda t: t.to(dtype).float())          # the kernels hold q, k (after the elu) and kv in the activation dtype, as the autocast graph does
    qq = r16(F.elu(q.float() + qb.view(1, -1, 1, 1)) + 1.0).view(B, heads, d, n)
    kk = r16(F.elu(k.float() + kb.view(1, -1, 1, 1)) + 1.0).view(B, heads, d, n)
    vv = v.float().view(B, heads, d, n)
    kvm = (kk @ vv.transpose(-1, -2)) / n
    num = qq.transpose(-1, -2) @ kvm
    den = qq.transpose(-1, -2) @ kk.mean(dim=-1, keepdim=True) + 1e-6
    ref = (num / den).transpose(-1, -2).reshape(B, dim, h, w) + pe
    tol = {torch.bfloat16: 1.5e-2, torch.float16: 3e-3, torch.float32: 2e-5}[dtype]
    out_pe = linattn_forward_pe(q, k, qb, kb, v, pw, pb, heads)
    assert rel_err(out_pe.float().cpu().numpy(), ref.cpu().numpy()) < tol
    out_qk = linattn_forward_qk(q, k, qb, kb, v, pe.to(dtype), heads)
    assert rel_err(out_qk.float().cpu().numpy(), ref.cpu().numpy()) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[8:-4])
def test_cuda_linear_attention_fp32_parity(path):
    """fp32 bar of the north star (1e-5) for the linear-attention kernel: fixture inputs/outputs of the unmodified reference"""
    from recnext_b200.recattn import linattn_forward
    from tests.helpers import TOL_FP32

    z, meta, sd = _load(path)
    m = _module(meta, sd)
    la = m.down[1]
    low = torch.from_numpy(z["low"])
    with torch.no_grad():
        qk_pre, pe = la.qk(low), la.pe(low)          # the two ConvNorms on the CPU in fp32 (reference arithmetic)
    out = linattn_forward(qk_pre.cuda(), low.cuda(), pe.cuda(), la.num_heads)
    assert out.dtype == torch.float32
    assert rel_err(out.cpu().numpy(), z["z"]) < TOL_FP32


@pytest.mark.gpu
def test_cuda_linear_attention_unsupported_head_dim_is_an_error():
    from recnext_b200.recattn import linattn_forward

    qk = torch.randn(1, 2 * 12, 4, 4, device="cuda").bfloat16()
    with pytest.raises(RuntimeError, match="head_dim"):
        linattn_forward(qk, torch.randn(1, 12, 4, 4, device="cuda").bfloat16(), None, 1)   # head_dim 12
