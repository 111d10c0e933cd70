"""Host-side RecNeXt-M mirror: state_dict layout and logits against the reference fixtures (tests/golden/model_logits.npz,
generated from the unmodified reference with oracle/detinit.py weights)."""
import os

import numpy as np
import pytest
import torch

from oracle.detinit import fill_state_dict_
from oracle.torch_ref import RefRecConv2d
from tests.helpers import GOLDEN, rel_err

Z = np.load(os.path.join(GOLDEN, "model_logits.npz"))


def _inputs(res):
    g = torch.Generator().manual_seed(1234)
    return torch.randn(2, 3, res, res, generator=g)


@pytest.mark.parametrize("variant", ["recnext_m0", "recnext_m3"])
def test_state_dict_keys_match_reference(variant):
    from recnext_b200.model import create_model

    net = create_model(variant)
    assert sorted(net.state_dict().keys()) == [str(k) for k in Z[f"{variant}_keys"]]
    assert sum(p.numel() for p in net.parameters()) == int(Z[f"{variant}_224_nparams"])


def test_host_model_logits_cpu_with_torch_token_mixer():
    """The surrounding model code (stem, BN, mlp, downsample, head) reproduces the reference logits on CPU."""
    from recnext_b200.model import create_model, replace_batchnorm

    net = create_model("recnext_m0", token_mixer=RefRecConv2d).eval()
    fill_state_dict_(net, seed=0)
    with torch.no_grad():
        y = net(_inputs(224)).numpy()
        assert rel_err(y, Z["recnext_m0_224_logits"]) < 1e-4
        replace_batchnorm(net)  # fused-BN eval model: same function
        assert rel_err(net(_inputs(224)).numpy(), Z["recnext_m0_224_logits"]) < 1e-4
    # after fusing, the un-fusable BatchNorm2d layers are exactly the block/downsample norms
    n_bn = sum(isinstance(m, torch.nn.BatchNorm2d) for m in net.modules())
    assert n_bn == sum((2, 2, 9, 1)) + 3


@pytest.mark.gpu
@pytest.mark.parametrize("variant,res", [("recnext_m0", 224), ("recnext_m3", 224), ("recnext_m0", 160)])
def test_full_model_logits_gpu(variant, res):
    """Full RecNeXt forward with the CUDA RecConv2d vs the reference's logits on identical inputs and weights."""
    from recnext_b200.model import create_model, replace_batchnorm

    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = create_model(variant).eval()
        fill_state_dict_(net, seed=0)
        net.cuda()
        with torch.no_grad():
            y = net(_inputs(res).cuda()).cpu().numpy()
            assert rel_err(y, Z[f"{variant}_{res}_logits"]) < 1e-4
            feats = net.forward_features(_inputs(res).cuda()).mean((2, 3)).cpu().numpy()
            assert rel_err(feats, Z[f"{variant}_{res}_feat_mean"]) < 1e-4
            replace_batchnorm(net)
            assert rel_err(net(_inputs(res).cuda()).cpu().numpy(), Z[f"{variant}_{res}_logits"]) < 1e-4
            with torch.autocast("cuda", dtype=torch.bfloat16):
                yb = net(_inputs(res).cuda()).float().cpu().numpy()
            assert rel_err(yb, Z[f"{variant}_{res}_logits"]) < 5e-2
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.gpu
def test_model_training_step_gradients_gpu():
    """One fwd+bwd of RecNeXt-M0 (training mode BN) with the CUDA token mixer vs the PyTorch restatement."""
    from recnext_b200.model import create_model

    torch.manual_seed(0)
    a = create_model("recnext_m0").cuda().train()
    b = create_model("recnext_m0", token_mixer=RefRecConv2d).cuda().train()
    b.load_state_dict(a.state_dict(), strict=True)
    x = torch.randn(4, 3, 128, 128, device="cuda")
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        la = a(x).logsumexp(1).mean(); la.backward()
        lb = b(x).logsumexp(1).mean(); lb.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert abs(la.item() - lb.item()) < 1e-4 * max(1.0, abs(lb.item()))
    # biases in front of a training-mode BatchNorm have a mathematically zero gradient (pure rounding noise), so the
    # error of every tensor is measured against max(|its reference gradient|, 1e-3 * largest gradient in the model)
    gmax = max(float(q.grad.abs().max()) for q in b.parameters())
    worst, worst_name = 0.0, None
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        diff = float((p.grad - q.grad).abs().max())
        e = diff / max(float(q.grad.abs().max()), 1e-3 * gmax)
        if e > worst:
            worst, worst_name = e, n
    assert worst < 2e-3, (worst, worst_name)


# ---- A-series (RecAttn2d token mixers; reference model/recattn.py) -------------------------------------------------
ZA = np.load(os.path.join(GOLDEN, "model_logits_a.npz"))


def test_a_series_state_dict_keys_match_reference():
    from recnext_b200.model import create_model, replace_batchnorm

    net = create_model("recnext_a0")
    assert sorted(net.state_dict().keys()) == [str(k) for k in ZA["recnext_a0_keys"]]
    assert sum(p.numel() for p in net.parameters()) == int(ZA["recnext_a0_224_nparams"])
    replace_batchnorm(net.eval())  # every ConvNorm / NormLinear folds: the A-series has no un-fusable BatchNorm (besides Downsample.norm)
    assert sum(isinstance(m, torch.nn.BatchNorm2d) for m in net.modules()) == 3


@pytest.mark.gpu
def test_a_series_full_model_logits_gpu():
    """RecNeXt-A0 inference (bf16 autocast, BatchNorm folded) with the CUDA RecAttn2d pieces vs the reference's fp32 logits."""
    from recnext_b200.model import create_model, replace_batchnorm

    net = create_model("recnext_a0").eval()
    fill_state_dict_(net, seed=0)
    net.cuda()
    x = _inputs(224).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y = net(x).float().cpu().numpy()
    assert rel_err(y, ZA["recnext_a0_224_logits"]) < 5e-2   # ~40 bf16 layers deep; per-op bar is 2e-2 (tests/test_recattn.py)
    replace_batchnorm(net)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = net(x).float().cpu().numpy()
    assert rel_err(y2, ZA["recnext_a0_224_logits"]) < 5e-2


@pytest.mark.gpu
def test_pipelined_inference_matches_plain_calls():
    """recnext_b200.infer.PipelinedInference (what bench.py times as e2e): same logits as calling the model batch by batch"""
    from recnext_b200.infer import PipelinedInference
    from recnext_b200.model import create_model, replace_batchnorm

    torch.manual_seed(0)
    net = create_model("recnext_m0").eval()
    fill_state_dict_(net, seed=0)
    replace_batchnorm(net)
    net.cuda()
    batches = [torch.randn(4, 3, 224, 224).bfloat16().pin_memory() for _ in range(5)]
    ref = []
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for b in batches:
            ref.append(net(b.cuda()).float().cpu())
    got = [y.float().clone() for y in PipelinedInference(net, torch.bfloat16).run(batches)]
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    # the same loop replaying one captured CUDA graph per staging buffer
    got = [y.float().clone() for y in PipelinedInference(net, torch.bfloat16, cuda_graph=True).run(batches)]
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["recnext_m0", "recnext_a0"])
def test_whole_model_is_cuda_graph_capturable(variant):
    """Every kernel of the eval path (stem, RecConv / RecAttn, linear attention, channel mixer, downsample) is stream-ordered, allocates
    nothing itself and never synchronises: a whole fused-BN model is captured in a CUDA graph and replayed on new data, bit-identically."""
    from recnext_b200.model import create_model

    torch.manual_seed(5)
    m = create_model(variant, num_classes=10).cuda().eval()
    m.fuse()
    x = torch.randn(4, 3, 64, 64, device="cuda").bfloat16()

    def fwd(inp):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return m(inp)

    ref1 = fwd(x)                                   # eager: also fills the packed-weight caches and the per-device kernel attributes
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = fwd(x)
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref1)
    x2 = torch.randn_like(x)
    ref2 = fwd(x2)
    x.copy_(x2)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref2)
