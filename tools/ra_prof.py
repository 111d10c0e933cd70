"""RecAttn2d pieces at the RecNeXt-A3 stage shapes (B = 256): python tools/ra_prof.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for C, H in ((64, 56), (128, 28), (256, 14), (512, 7)):
    x = torch.randn(256, C, H, H, device="cuda").bfloat16()
    w = torch.randn(C, 1, 5, 5, device="cuda") * 0.2; b = 0.1 * torch.randn(C, device="cuda")
    low = R.recattn_down_forward(x, w, b)
    z = torch.randn_like(low)
    R.recattn_up_forward(x, z, w, b, "nearest")
    td = timeit(lambda: R.recattn_down_forward(x, w, b)); tu = timeit(lambda: R.recattn_up_forward(x, z, w, b, "nearest"))
    nb = x.numel() * 2
    print(f"recattn [256,{C},{H},{H}]: down {td:.4f} ms ({1.25 * nb / td / 1e6:.0f} GB/s)   up {tu:.4f} ms ({2.25 * nb / tu / 1e6:.0f} GB/s)")
