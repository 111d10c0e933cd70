"""Fused channel mixer (recnext_ffn_forward; reference model/recnext.py:125-131,153,157-158): the kernel against the
PyTorch graph it replaces, the host-side BatchNorm fold, and the block-level switch in recnext_b200.model."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import TOL_BF16, rel_err


def _block(C, stage=0):
    from oracle.torch_ref import RefRecConv2d
    from recnext_b200.model import MetaNeXtBlock, replace_batchnorm

    torch.manual_seed(C)
    blk = MetaNeXtBlock(C, 2, stage=stage, token_mixer=RefRecConv2d)
    g = torch.Generator().manual_seed(1)
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.3 * torch.randn(m.num_features, generator=g)); m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.7 + 0.6 * torch.rand(m.num_features, generator=g)); m.bias.data.copy_(0.2 * torch.randn(m.num_features, generator=g))
    blk.eval()
    return blk, replace_batchnorm


def test_bn_fold_into_first_conv_cpu():
    """host logic: W1 (s*y + t) + b1 == (W1 diag(s)) y + (b1 + W1 t), so the folded parameters reproduce norm -> fc1"""
    blk, replace_batchnorm = _block(32)
    y = torch.randn(2, 32, 8, 8)
    with torch.no_grad():
        ref = blk.channel_mixer[0](blk.norm(y))            # ConvNorm(BN(y))
        replace_batchnorm(blk)
        import recnext_b200.model as M
        w1, b1, w2, b2 = M._fold_mlp(blk.channel_mixer[0], blk.channel_mixer[2], torch.float32, blk.norm)
        got = F.conv2d(y, w1.view(64, 32, 1, 1), b1)
        assert rel_err(got.numpy(), ref.numpy()) < 1e-5
        assert tuple(w2.shape) == (32, 64) and tuple(b2.shape) == (32,)
    assert not blk._ffn_eligible(y)                        # CPU tensors never take the kernel (and nothing falls back silently)
    blk.train()
    assert blk._ffn_cache is None                          # cached fold is dropped when the statistics may change
    # the cache key carries the version counter of every source tensor: load_state_dict / in-place updates invalidate it
    k0 = M._param_key(blk.channel_mixer[0], blk.channel_mixer[2], blk.norm)
    blk.load_state_dict(blk.state_dict())
    assert M._param_key(blk.channel_mixer[0], blk.channel_mixer[2], blk.norm) != k0
    # with gradients enabled and trainable parameters the inference kernels step aside (they run outside autograd)
    assert M._needs_autograd(y, blk)
    with torch.no_grad():
        assert not M._needs_autograd(y, blk)


def test_ffn_forward_has_no_cpu_path():
    from recnext_b200.model import ffn_forward

    with pytest.raises(RuntimeError, match="CUDA"):
        ffn_forward(torch.randn(1, 16, 4, 4), torch.randn(1, 16, 4, 4), torch.randn(32, 16), torch.randn(32), torch.randn(16, 32), torch.randn(16))


# (B, C, H, W, hidden): every RecNeXt width of the M / A series (reference model/recnext.py:369-406, model/recattn.py:382-419; hidden = 2C
# or 1.875C) at its stage's plane size, plus ragged cases: pixel tiles that straddle images, HW % 8 != 0 (8- and 2-byte accesses),
# C % 16 != 0 (zero-padded K), hidden % 128 != 0, more tiles than SMs
_FFN_CASES = [(4, 64, 56, 56, 128), (4, 128, 28, 28, 256), (3, 256, 14, 14, 512), (3, 512, 7, 7, 1024), (3, 80, 28, 28, 160), (2, 160, 14, 14, 320),
              (5, 320, 14, 14, 640), (3, 640, 7, 7, 1280), (2, 40, 56, 56, 80), (2, 64, 56, 56, 120), (2, 128, 28, 28, 240), (2, 32, 6, 6, 64),
              (2, 48, 10, 18, 96), (3, 24, 5, 7, 40), (64, 64, 56, 56, 128), (2, 256, 50, 84, 512),
              (7, 64, 7, 7, 128), (2, 152, 9, 13, 304), (37, 320, 7, 7, 600)]   # odd planes: pixel-per-lane loaders / writers, partial channel tiles


@pytest.mark.gpu
@pytest.mark.parametrize("case", _FFN_CASES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_ffn_kernel_vs_torch(case, dtype):
    from recnext_b200.model import ffn_forward

    B, C, H, W, hid = case
    shape = (B, C, H, W)
    torch.manual_seed(C + H)
    y = torch.randn(shape, device="cuda").to(dtype); x = torch.randn(shape, device="cuda").to(dtype)
    w1 = (torch.randn(hid, C, device="cuda") * C ** -0.5).to(dtype); w2 = (torch.randn(C, hid, device="cuda") * hid ** -0.5).to(dtype)
    b1 = 0.1 * torch.randn(hid, device="cuda"); b2 = 0.1 * torch.randn(C, device="cuda")
    out = ffn_forward(y, x, w1, b1, w2, b2)
    ref = x.float() + F.conv2d(F.gelu(F.conv2d(y.float(), w1.float().view(hid, C, 1, 1), b1)), w2.float().view(C, hid, 1, 1), b2)
    assert rel_err(out.float().cpu().numpy(), ref.cpu().numpy()) < (TOL_BF16 if dtype == torch.bfloat16 else 4e-3)
    # the kernel is deterministic and every pixel is independent: a batch slice gives bit-identical rows
    assert torch.equal(out, ffn_forward(y, x, w1, b1, w2, b2))
    if B > 1:
        assert torch.equal(out[1:], ffn_forward(y[1:].contiguous(), x[1:].contiguous(), w1, b1, w2, b2))


@pytest.mark.gpu
def test_ffn_unsupported_shapes_are_errors():
    from recnext_b200.model import ffn_forward

    y = torch.randn(1, 36, 7, 7, device="cuda").bfloat16()   # C % 8 != 0
    with pytest.raises(RuntimeError, match="ffn_pack"):
        ffn_forward(y, y, torch.randn(72, 36, device="cuda").bfloat16(), torch.randn(72, device="cuda"), torch.randn(36, 72, device="cuda").bfloat16(),
                    torch.randn(36, device="cuda"))


@pytest.mark.gpu
def test_block_with_fused_ffn_matches_library_path(monkeypatch):
    """MetaNeXtBlock in eval mode with folded ConvNorms: fused kernel path vs the PyTorch module graph (same parameters)"""
    import recnext_b200.model as M

    blk, replace_batchnorm = _block(64)
    replace_batchnorm(blk)
    blk.cuda()
    x = torch.randn(4, 64, 56, 56, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert blk._ffn_eligible(x)
        y_fused = blk(x)
        monkeypatch.setattr(M, "FUSED_FFN", False)
        assert not blk._ffn_eligible(x)
        y_lib = blk(x)
    ref32 = None
    with torch.no_grad():
        ref32 = x + blk.channel_mixer(blk.norm(blk.token_mixer(x)))
    assert y_fused.dtype == torch.bfloat16
    assert rel_err(y_fused.float().cpu().numpy(), ref32.cpu().numpy()) < TOL_BF16
    assert rel_err(y_lib.float().cpu().numpy(), ref32.cpu().numpy()) < TOL_BF16


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 64, 56, 56), (3, 128, 28, 28), (3, 256, 14, 14), (2, 40, 56, 56), (2, 6, 9, 13), (1, 3, 25, 42), (2, 4, 7, 7), (1, 8, 112, 112), (2, 16, 30, 20),
                                   (40, 7, 14, 14), (1, 2, 1, 1), (1, 5, 2, 3)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
def test_dwdown_kernel_vs_torch(shape, dtype):
    """Downsample token mixer (depthwise 7x7 stride 2, multiplier 2, + bias; reference model/recnext.py:137-138) vs F.conv2d"""
    from recnext_b200.model import dwdown_forward

    B, C, H, W = shape
    torch.manual_seed(C + H)
    x = torch.randn(shape, device="cuda").to(dtype)
    w = torch.randn(2 * C, 1, 7, 7, device="cuda") / 7.0
    b = 0.1 * torch.randn(2 * C, device="cuda")
    out = dwdown_forward(x, w, b)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=3, groups=C)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert out.shape == ref.shape
    assert rel_err(out.float().cpu().numpy(), ref.cpu().numpy()) < {torch.float32: 1e-5, torch.bfloat16: TOL_BF16, torch.float16: 4e-3}[dtype]


@pytest.mark.gpu
def test_downsample_block_fused_matches_library_path(monkeypatch):
    import recnext_b200.model as M

    torch.manual_seed(3)
    blk = M.Downsample(64, 2)
    g = torch.Generator().manual_seed(1)
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.3 * torch.randn(m.num_features, generator=g)); m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.7 + 0.6 * torch.rand(m.num_features, generator=g)); m.bias.data.copy_(0.2 * torch.randn(m.num_features, generator=g))
    blk.eval()
    M.replace_batchnorm(blk)
    blk.cuda()
    x = torch.randn(4, 64, 56, 56, device="cuda")
    with torch.no_grad():
        ref = blk.norm(blk.token_mixer(x))
        ref = ref + blk.channel_mixer(ref)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = blk(x)
            monkeypatch.setattr(M, "FUSED_FFN", False)
            y_lib = blk(x)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (4, 128, 28, 28)
    assert rel_err(y.float().cpu().numpy(), ref.cpu().numpy()) < TOL_BF16
    assert rel_err(y_lib.float().cpu().numpy(), ref.cpu().numpy()) < TOL_BF16


@pytest.mark.gpu
def test_dwdown_full_size_batch_slices_bit_identical():
    """BASELINE configs[1] stage borders at batch 256: plane groups never mix (an image computed inside the batch equals the image alone)"""
    from recnext_b200.model import dwdown_forward

    for C, H in ((64, 56), (128, 28), (256, 14)):
        torch.manual_seed(C)
        x = torch.randn(256, C, H, H, device="cuda").bfloat16()
        w = torch.randn(2 * C, 1, 7, 7, device="cuda") / 7.0
        b = 0.1 * torch.randn(2 * C, device="cuda")
        full = dwdown_forward(x, w, b)
        for i in (0, 77, 255):
            assert torch.equal(dwdown_forward(x[i:i + 1].contiguous(), w, b)[0], full[i])
        ref = F.conv2d(x[:2].float(), w, b, stride=2, padding=3, groups=C)
        assert rel_err(full[:2].float().cpu().numpy(), ref.cpu().numpy()) < TOL_BF16
