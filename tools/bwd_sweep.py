"""Backward timing for one shape under the plan overrides in the environment.   python tools/bwd_sweep.py B C H W L"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R
B, C, H, W, L = (int(v) for v in sys.argv[1:6])
m = R.RecConv2d(C, level=L).cuda()
ws = [w.detach() for w in m._param_lists()[0]]
x = torch.randn(B, C, H, W, device="cuda").bfloat16(); gy = torch.randn_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(7):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); R.recconv_backward(x, gy, ws, None, 5, L, "bilinear"); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
desc = R.plan_describe((B, C, H, W), 5, L, "bilinear", torch.bfloat16, False, True)
print(f"{sorted(ts)[3]:.3f} ms  env {dict((k, v) for k, v in os.environ.items() if k.startswith('RECNEXT_MB'))}  {desc[4:15]} {desc[desc.find('planes/batch'):desc.find('grid')]}")
