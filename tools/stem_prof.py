"""Fused stem against the library graph: python tools/stem_prof.py [B] [C] [H]"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recnext_b200.model import RecNextStem, replace_batchnorm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = int(sys.argv[3]) if len(sys.argv) > 3 else 224
torch.manual_seed(0)
m = RecNextStem(3, C).eval().cuda()
replace_batchnorm(m)
x = torch.randn(B, 3, H, H, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
import recnext_b200.model as M
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3): m(x)
    t_fused = timeit(lambda: m(x))
    M.FUSED_STEM = False
    for _ in range(3): m(x)
    t_lib = timeit(lambda: m(x))
nbytes = (x.numel() + B * C * (H // 4) ** 2) * 2
print(f"stem [{B},3,{H},{H}] -> {C}: fused {t_fused:.4f} ms ({nbytes / t_fused / 1e6:.0f} GB/s of in+out)   library graph {t_lib:.4f} ms")
