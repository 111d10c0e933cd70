"""Downsample token mixer (depthwise 7x7 stride 2, multiplier 2) at the RecNeXt-M3 stage borders: python tools/dw_prof.py [B]"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recnext_b200.model import dwdown_forward
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for C, H in ((64, 56), (128, 28), (256, 14), (80, 56), (160, 28), (320, 14)):
    x = torch.randn(B, C, H, H, device="cuda").bfloat16()
    w = torch.randn(2 * C, 1, 7, 7, device="cuda") / 7.0
    b = 0.1 * torch.randn(2 * C, device="cuda")
    for _ in range(2): dwdown_forward(x, w, b)
    t = timeit(lambda: dwdown_forward(x, w, b))
    nbytes = x.numel() * 2 + B * 2 * C * (H // 2) ** 2 * 2
    t_lib = timeit(lambda: F.conv2d(x, w.bfloat16(), b.bfloat16(), stride=2, padding=3, groups=C))
    print(f"dwdown [{B},{C},{H},{H}]: {t:.4f} ms ({nbytes / t / 1e6:.0f} GB/s of in+out)   F.conv2d {t_lib:.4f} ms")
