"""Per-stage clock64 stamps of team 0 / CTA 0 of the tensor-core backward (RECNEXT_PROF=1).  python tools/bwd_prof.py B C H W L"""
import ctypes, os, sys
import numpy as np, torch
os.environ["RECNEXT_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R
from recnext_b200 import _native as N
B, C, H, W, L = (int(v) for v in sys.argv[1:6])
m = R.RecConv2d(C, level=L).cuda()
ws = [w.detach() for w in m._param_lists()[0]]
x = torch.randn(B, C, H, W, device="cuda").bfloat16(); gy = torch.randn_like(x)
for _ in range(2): R.recconv_backward(x, gy, ws, None, 5, L, "bilinear")
torch.cuda.synchronize()
buf = np.zeros(4096, dtype=np.int64)
lib = N.lib(); lib.recnext_debug_prof.restype = ctypes.c_int; lib.recnext_debug_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.recnext_debug_prof(buf.ctypes.data, 4096)
st = buf[buf != 0]
d = np.diff(st)
n = 3 + 3 * L + 3 + 3 * L + 2 * L   # stamps per plane
print(R.plan_describe((B, C, H, W), 5, L, "bilinear", torch.bfloat16, False, True))
print("stamps per plane", n, "total stamps", len(st))
names = ["repack x"] + [f"down {l}" for l in range(1, L + 1)] + ["copy X->S"] + sum([[f"conv lvl {l}", f"up-add {l}"] for l in range(L, 0, -1)], []) + ["zero+repack gy", "wgrad L0", "dgrad L0"] \
        + sum([[f"gather {l}", f"wgrad {l}", f"dgrad {l}"] for l in range(1, L + 1)], []) + sum([[f"wgrad_s2 {l}", f"down^T {l}"] for l in range(L, 0, -1)], []) + ["(next plane)"]
per = len(names)
for p in range(1, 3):
    seg = d[p * per:(p + 1) * per]
    print("plane", p, "total", int(seg.sum()))
    for nm, v in zip(names, seg): print(f"   {nm:18s} {int(v):8d}")
