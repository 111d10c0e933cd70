"""Fused channel-mixer kernel vs PyTorch (fp32 and bf16 module graph) + timing.   python tools/ffn_check.py"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recnext_b200.model import ffn_forward
dev = "cuda"
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
for (B, C, H) in [(4, 64, 56), (4, 128, 28), (4, 256, 14), (3, 80, 28), (2, 160, 14), (2, 32, 6), (256, 64, 56), (256, 128, 28), (256, 256, 14), (256, 80, 56), (256, 160, 28), (256, 320, 14)]:
    torch.manual_seed(C + H)
    hid = 2 * C
    y = torch.randn(B, C, H, H, device=dev).bfloat16(); x = torch.randn(B, C, H, H, device=dev).bfloat16()
    w1 = (torch.randn(hid, C, device=dev) * C ** -0.5).bfloat16(); w2 = (torch.randn(C, hid, device=dev) * hid ** -0.5).bfloat16()
    b1 = torch.randn(hid, device=dev) * 0.1; b2 = torch.randn(C, device=dev) * 0.1
    out = ffn_forward(y, x, w1, b1, w2, b2)
    torch.cuda.synchronize()
    n = min(B, 8)
    ref = x[:n].float() + F.conv2d(F.gelu(F.conv2d(y[:n].float(), w1.float().view(hid, C, 1, 1), b1)), w2.float().view(C, hid, 1, 1), b2)
    msg = f"[{B},{C},{H},{H}] hid {hid}: rel err vs fp32 {rel(out[:n], ref):.2e}"
    if B >= 64:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(5):
            flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ffn_forward(y, x, w1, b1, w2, b2); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = sorted(ts)[2]
        gb = 3 * y.numel() * 2 / 1e9
        msg += f"   {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s of 3Ne"
        # the library path the fused kernel replaces: BatchNorm(eval) -> 1x1 conv -> GELU -> 1x1 conv -> + x  (bf16, cuDNN)
        bn = torch.nn.BatchNorm2d(C).to(dev).eval().bfloat16()
        c1 = torch.nn.Conv2d(C, hid, 1).to(dev).bfloat16(); c2 = torch.nn.Conv2d(hid, C, 1).to(dev).bfloat16()
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            f = lambda: x + c2(F.gelu(c1(bn(y))))
            for _ in range(3): f()
            ts = []
            for _ in range(5):
                flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        msg += f"   | library path {sorted(ts)[2]:.3f} ms"
    print(msg)
