"""Fused stem (recnext_stem_forward; reference model/recnext.py:139-146): conv 3x3 s2 -> GELU -> conv 3x3 s2 as one kernel against the
PyTorch graph it replaces, the host-side packing, and the module-level switch in recnext_b200.model."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err


def _stem(C, seed=0):
    from recnext_b200.model import RecNextStem, replace_batchnorm

    torch.manual_seed(seed)
    m = RecNextStem(3, C)
    g = torch.Generator().manual_seed(1)
    for bn in m.modules():
        if isinstance(bn, torch.nn.BatchNorm2d):
            bn.running_mean.copy_(0.3 * torch.randn(bn.num_features, generator=g)); bn.running_var.copy_(0.5 + torch.rand(bn.num_features, generator=g))
            bn.weight.data.copy_(0.7 + 0.6 * torch.rand(bn.num_features, generator=g)); bn.bias.data.copy_(0.2 * torch.randn(bn.num_features, generator=g))
    m.eval()
    return m, replace_batchnorm


def _reference(x, conv1, conv2, dtype):
    """the reference's autocast graph with fp32 accumulation: conv -> 16-bit -> GELU (erf) -> 16-bit -> conv -> 16-bit"""
    r = lambda t: t.to(dtype).float()
    h = r(F.conv2d(r(x), r(conv1.weight), conv1.bias.float(), stride=2, padding=1))
    h = r(F.gelu(h))
    return r(F.conv2d(h, r(conv2.weight), conv2.bias.float(), stride=2, padding=1))


def test_stem_pack_layout_cpu():
    """host logic: the packed operands are the conv weights in the kernel's K order, zero padded"""
    from recnext_b200.model import stem_pack

    m, replace_batchnorm = _stem(40)
    replace_batchnorm(m)
    c1, c2 = m.stem[0], m.stem[2]
    w1p, b1p, w2p, b2p, C1, C2 = stem_pack(c1, c2, torch.float32)
    assert (C1, C2) == (20, 40) and tuple(w1p.shape) == (32, 32) and tuple(w2p.shape) == (48, 9 * 32)
    assert torch.equal(w1p[7, 1 * 9 + 2 * 3 + 1], c1.weight[7, 1, 2, 1]) and float(w1p[:, 27:].abs().sum()) == 0 and float(w1p[20:].abs().sum()) == 0
    assert torch.equal(w2p[33, (1 * 3 + 2) * 32 + 11], c2.weight[33, 11, 1, 2]) and float(w2p[40:].abs().sum()) == 0
    assert float(w2p.view(48, 9, 32)[:, :, 20:].abs().sum()) == 0 and torch.equal(b2p[:40], c2.bias) and float(b1p[20:].abs().sum()) == 0
    x = torch.randn(1, 3, 32, 32)
    assert not m._fused_ok(x)                       # CPU tensors never take the kernel (and nothing falls back silently)


def test_stem_forward_has_no_cpu_path():
    from recnext_b200.model import stem_forward

    with pytest.raises(RuntimeError, match="CUDA"):
        stem_forward(torch.randn(1, 3, 8, 8).bfloat16(), None, None, None, None, 20, 40)


# every stem width of the M / A series (40 .. 80) and ragged sizes: maps that are not multiples of the 8 x 8 tile, odd sizes, a single pixel
_CASES = [(2, 64, 224, 224), (3, 40, 64, 64), (2, 48, 96, 80), (2, 56, 70, 50), (2, 80, 64, 96), (1, 64, 33, 47), (1, 64, 200, 336), (2, 40, 1, 1),
          (1, 72, 40, 40)]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,C,H,W", _CASES)
def test_stem_kernel_matches_graph(B, C, H, W, dtype):
    from recnext_b200.model import stem_forward, stem_pack

    m, replace_batchnorm = _stem(C, seed=C + H)
    replace_batchnorm(m)
    m = m.cuda()
    x = torch.randn(B, 3, H, W, device="cuda")
    with torch.no_grad():
        got = stem_forward(x.to(dtype), *stem_pack(m.stem[0], m.stem[2], dtype)).float()
        ref = _reference(x, m.stem[0], m.stem[2], dtype)
        ref32 = F.conv2d(F.gelu(F.conv2d(x, m.stem[0].weight, m.stem[0].bias, stride=2, padding=1)), m.stem[2].weight, m.stem[2].bias, stride=2, padding=1)
    assert got.shape == ref.shape
    g, r, r32 = got.cpu().numpy(), ref.cpu().numpy(), ref32.cpu().numpy()
    # The 16-bit graph is itself a rounded version of the fp32 graph (measured: 3.4e-3 relative in bf16, 4.5e-4 in fp16).  The kernel rounds
    # at the same places, so (a) it is as close to the fp32 graph as the 16-bit graph is, and (b) the two differ by about two such noises
    # (a rounding flip of the intermediate moves an output by one 16-bit ulp); the tanh fit of GELU adds 2.7e-4 absolute.
    noise = rel_err(r, r32)
    assert rel_err(g, r32) < 1.3 * noise + 2e-4
    assert rel_err(g, r) < 2.0 * noise + 2e-4
    assert float((got - ref).abs().max()) < 0.08 * float(ref.abs().max()) + 1e-2


@pytest.mark.gpu
def test_stem_module_switch():
    """RecNextStem takes the kernel in eval mode with folded ConvNorms under autocast, and the PyTorch graph otherwise"""
    import recnext_b200.model as M

    m, replace_batchnorm = _stem(64)
    m = m.cuda()
    x = torch.randn(2, 3, 64, 64, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert not m._fused_ok(x)                   # ConvNorms not folded yet
        ref = m(x)
        replace_batchnorm(m)
        assert m._fused_ok(x)
        got = m(x)
        assert got.dtype == torch.bfloat16 and got.shape == ref.shape
        assert rel_err(got.float().cpu().numpy(), ref.float().cpu().numpy()) < 1.5e-2
        k0 = m._stem_cache[0]
        m(x)
        assert m._stem_cache[0] == k0               # packed once
        m.load_state_dict(m.state_dict())
        m(x)
        assert m._stem_cache[0] != k0               # ... and again after the parameters changed
    m.train()
    assert m._stem_cache is None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert not m._fused_ok(x)                   # training mode / gradients wanted: the differentiable graph runs


@pytest.mark.gpu
def test_full_size_batch_slices_bit_identical():
    """BASELINE configs[1] size (256 x 3 x 224 x 224): every image is computed exactly as it is alone (tiles never mix images; no atomics), and
    the interior agrees with the library graph on a sample of images"""
    from recnext_b200.model import stem_forward, stem_pack

    m, replace_batchnorm = _stem(64)
    replace_batchnorm(m)
    m = m.cuda()
    pk = stem_pack(m.stem[0], m.stem[2], torch.bfloat16)
    x = torch.randn(256, 3, 224, 224, device="cuda").bfloat16()
    with torch.no_grad():
        full = stem_forward(x, *pk)
        for b in (0, 101, 255):
            assert torch.equal(stem_forward(x[b:b + 1].contiguous(), *pk)[0], full[b])
        ref = _reference(x[:4].float(), m.stem[0], m.stem[2], torch.bfloat16)
    assert tuple(full.shape) == (256, 64, 56, 56)
    assert rel_err(full[:4].float().cpu().numpy(), ref.cpu().numpy()) < 8e-3
