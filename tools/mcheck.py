"""Debug check of the tensor-core forward against PyTorch (fp32 and bf16-autocast semantics) on the GPU.
    python tools/mcheck.py            (RECNEXT_PATH=fma python tools/mcheck.py for the FMA kernels)"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import recnext_b200 as R
from recnext_b200 import _native

def ref(x, ws, bs, L, mode, dt):
    C = x.shape[1]
    cast = lambda t: None if t is None else t.to(dt)
    feats, cur = [], x.to(dt)
    for _ in range(L):
        size = cur.shape[2:]
        cur = F.conv2d(cur, cast(ws[0]), cast(bs[0]) if bs else None, stride=2, padding=2, groups=C)
        feats.append((cur, size))
    up = 0
    for j, (f, size) in enumerate(reversed(feats)):
        up = F.interpolate(F.conv2d(f + up, cast(ws[1 + j]), cast(bs[1 + j]) if bs else None, padding=2, groups=C), size=size, mode=mode)
    return F.conv2d(x.to(dt) + up, cast(ws[1 + L]), cast(bs[1 + L]) if bs else None, padding=2, groups=C)

CASES = [(3, 64, 56, 56, 4, "bilinear", False), (3, 128, 28, 28, 3, "bilinear", False), (5, 256, 14, 14, 2, "bilinear", True),
         (5, 512, 7, 7, 1, "bilinear", False), (2, 80, 56, 56, 4, "nearest", True), (2, 256, 50, 84, 2, "bilinear", False),
         (2, 512, 25, 42, 1, "bilinear", True), (1, 7, 25, 21, 3, "bilinear", True), (1, 1, 1, 1, 2, "bilinear", True),
         (4, 6, 2, 3, 1, "nearest", False), (7, 9, 8, 8, 0, "bilinear", True), (2, 128, 100, 168, 3, "bilinear", False),
         (2, 128, 100, 167, 3, "bilinear", False), (64, 64, 56, 56, 4, "bilinear", False), (1, 16, 96, 96, 4, "bilinear", False),
         (2, 64, 200, 336, 4, "bilinear", False), (1, 8, 200, 334, 4, "bilinear", True), (1, 4, 200, 336, 4, "nearest", False)]
dev = "cuda"
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
bad = 0
for dt in (torch.bfloat16, torch.float16):
    for (B, C, H, W, L, mode, bias) in CASES:
        torch.manual_seed(B * 1000 + C + H)
        x = torch.randn(B, C, H, W, device=dev).to(dt)
        ws = [torch.empty(C, 1, 5, 5, device=dev).uniform_(-0.2, 0.2) for _ in range(L + 2)]
        bs = [torch.empty(C, device=dev).uniform_(-0.2, 0.2) for _ in range(L + 2)] if bias else None
        desc = R.recconv.plan_describe((B, C, H, W), 5, L, mode, dt, bias, False) if (B, C) == (3, 64) or H >= 96 else ""
        if dt == torch.float16 and H >= 200: continue
        try:
            y = R.recconv_forward(x, ws, bs, 5, L, mode)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"{str(dt)[6:]:9s} {(B,C,H,W,L,mode,bias)}: ERROR {e}"); bad += 1; continue
        y32 = ref(x.float(), ws, bs, L, mode, torch.float32)
        ylo = ref(x, ws, bs, L, mode, dt)
        e32, elo, eref = rel(y, y32), rel(y, ylo), rel(ylo, y32)
        tol = 2e-2 if dt == torch.bfloat16 else 4e-3
        flag = "" if (e32 < tol and elo < tol) else "   <<<<<< FAIL"
        bad += bool(flag)
        print(f"{str(dt)[6:]:9s} {(B,C,H,W,L,mode,bias)}: vs fp32 {e32:.2e}  vs torch-{str(dt)[6:]} {elo:.2e}  (torch-lowp vs fp32 {eref:.2e}){flag}  {desc}")
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
