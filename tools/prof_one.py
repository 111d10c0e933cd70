"""One RecConv forward + backward launch per iteration, for ncu.  python tools/prof_one.py B C H W L [dtype] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R  # noqa: E402

B, C, H, W, L = (int(v) for v in sys.argv[1:6])
dt = {"bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[6] if len(sys.argv) > 6 else "bf16"]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 2
what = sys.argv[8] if len(sys.argv) > 8 else "both"
m = R.RecConv2d(C, level=L).cuda()
ws = [w.detach() for w in m._param_lists()[0]]
x = torch.randn(B, C, H, W, device="cuda").to(dt)
gy = torch.randn_like(x)
for _ in range(iters):
    if what in ("both", "fwd"):
        R.recconv_forward(x, ws, None, 5, L, "bilinear")
    if what in ("both", "bwd"):
        R.recconv_backward(x, gy, ws, None, 5, L, "bilinear")
torch.cuda.synchronize()
