"""Linear-attention kernel at the RecNeXt-A3 stage shapes (B = 256): python tools/la_prof.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recnext_b200.recattn import linattn_forward_pe
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for dim, heads, h in ((64, 2, 28), (128, 4, 14), (256, 8, 7), (512, 16, 4)):
    q = torch.randn(256, dim, h, h, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
    qb = torch.randn(dim, device="cuda"); kb = torch.randn(dim, device="cuda"); pw = torch.randn(dim, 1, 3, 3, device="cuda"); pb = torch.randn(dim, device="cuda")
    f = lambda: linattn_forward_pe(q, k, qb, kb, v, pw, pb, heads)
    f(); f()
    print(f"linattn [256,{dim},{h},{h}] heads {heads}: {timeit(f):.4f} ms  ({4 * q.numel() * 2 / timeit(f) / 1e6:.0f} GB/s of q,k,v,out)")
