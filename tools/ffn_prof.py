"""One fused-FFN launch for ncu.  python tools/ffn_prof.py B C H"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recnext_b200.model import ffn_forward
B, C, H = (int(v) for v in sys.argv[1:4])
hid = 2 * C
y = torch.randn(B, C, H, H, device="cuda").bfloat16(); x = torch.randn_like(y)
w1 = (torch.randn(hid, C, device="cuda") * C ** -0.5).bfloat16(); w2 = (torch.randn(C, hid, device="cuda") * hid ** -0.5).bfloat16()
b1 = torch.randn(hid, device="cuda") * 0.1; b2 = torch.randn(C, device="cuda") * 0.1
for _ in range(2):
    ffn_forward(y, x, w1, b1, w2, b2)
torch.cuda.synchronize()
