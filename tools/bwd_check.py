"""Accuracy of the RecConv backward against the fp32 PyTorch restatement (bf16 inputs), per stage shape.   python tools/bwd_check.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R
from oracle.torch_ref import recconv_reference
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
for shape, L, mode, bias in [((8, 64, 56, 56), 4, "bilinear", False), ((8, 128, 28, 28), 3, "bilinear", True), ((8, 256, 14, 14), 2, "nearest", False),
                             ((2, 128, 100, 168), 3, "bilinear", False), ((2, 16, 25, 42), 1, "bilinear", True), ((2, 8, 100, 167), 3, "bilinear", False)]:
    torch.manual_seed(0)
    C = shape[1]
    m = R.RecConv2d(C, level=L, mode=mode, bias=bias).cuda()
    ws, bs = m._param_lists()
    ws = [w.detach() for w in ws]; bs = [b.detach() for b in bs] if bias else None
    x = torch.randn(shape, device="cuda").bfloat16(); gy = torch.randn(shape, device="cuda").bfloat16()
    gx, gw, gb = R.recconv_backward(x, gy, ws, bs, 5, L, mode)
    xr = x.float().requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    br = [b.clone().requires_grad_(True) for b in bs] if bias else None
    torch.backends.cudnn.allow_tf32 = False
    y = recconv_reference(xr, wr[0], wr[1:], br[0] if bias else None, br[1:] if bias else None, mode)
    y.backward(gy.float())
    errs = [rel(gx.float(), xr.grad)] + [rel(gw[j].view(C, 1, 5, 5), wr[j].grad) for j in range(L + 2)]
    if bias:
        errs += [rel(gb[j], br[j].grad) for j in range(L + 2)]
    print(shape, L, mode, "bias" if bias else "", R.plan_describe(shape, 5, L, mode, torch.bfloat16, bias, True)[:18], " ".join(f"{e:.1e}" for e in errs))
