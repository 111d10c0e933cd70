"""torch.library registration of the RecConv kernels (reference publish.py:32-38, fuse_eval.py:48, export_coreml.py:30):
schema + fake (meta) implementations on CPU, opcheck / jit.trace / torch.compile on the GPU."""
import pytest
import torch


def _params(C, L, k=5, bias=False, device="cpu"):
    ws = [torch.randn(C, 1, k, k, device=device) * 0.2 for _ in range(L + 2)]
    bs = [torch.randn(C, device=device) * 0.1 for _ in range(L + 2)] if bias else []
    return ws, bs


def test_operator_is_registered_with_fake_impl():
    import recnext_b200.ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode

    assert "recconv" in dir(torch.ops.recnext) or hasattr(torch.ops.recnext, "recconv")
    schema = str(torch.ops.recnext.recconv.default._schema)
    assert "Tensor[] weights" in schema and "str mode" in schema
    with FakeTensorMode():   # shape propagation needs neither a GPU nor the native library
        x = torch.empty(2, 8, 14, 14, device="cuda", dtype=torch.bfloat16)
        ws = [torch.empty(8, 1, 5, 5, device="cuda") for _ in range(4)]
        y = torch.ops.recnext.recconv(x, ws, [], 5, 2, "bilinear")
        assert y.shape == x.shape and y.dtype == x.dtype
        gx, gw, gb = torch.ops.recnext.recconv_backward(x, x, ws, [], 5, 2, "bilinear")
        assert gx.shape == x.shape and tuple(gw.shape) == (4, 8, 25) and gw.dtype == torch.float32 and gb.numel() == 0


def test_module_on_cpu_still_raises():
    import recnext_b200 as R

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        R.RecConv2d(4, level=1)(torch.randn(1, 4, 8, 8))


@pytest.mark.gpu
def test_opcheck():
    import recnext_b200.ops  # noqa: F401

    x = torch.randn(2, 6, 14, 14, device="cuda", requires_grad=True)
    ws, bs = _params(6, 2, bias=True, device="cuda")
    for t in ws + bs:
        t.requires_grad_(True)
    torch.library.opcheck(torch.ops.recnext.recconv.default, (x, ws, bs, 5, 2, "bilinear"),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


@pytest.mark.gpu
def test_jit_trace_and_compile_see_one_operator():
    import recnext_b200 as R

    torch.manual_seed(0)
    m = R.RecConv2d(16, level=2, bias=True).cuda()
    x = torch.randn(2, 16, 14, 14, device="cuda")
    y = m(x)
    traced = torch.jit.trace(m, x)
    assert "recnext::recconv" in str(traced.graph)
    assert torch.equal(traced(x), y)
    cm = torch.compile(m, backend="aot_eager", fullgraph=True)
    xc = x.clone().requires_grad_(True)
    yc = cm(xc)
    assert torch.equal(yc, y)
    gy = torch.randn_like(y)
    yc.backward(gy)
    xe = x.clone().requires_grad_(True)
    for p in m.parameters():
        p.grad = None
    m(xe).backward(gy)
    assert torch.equal(xc.grad, xe.grad)
