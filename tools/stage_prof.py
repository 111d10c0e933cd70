"""Per-stage cycle counts of team 0 / CTA 0 (RECNEXT_PROF=1).  python tools/stage_prof.py B C H W L [dtype] [fwd|bwd]"""
import ctypes, os, sys
os.environ["RECNEXT_PROF"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R
from recnext_b200 import _native as N
B, C, H, W, L = (int(v) for v in sys.argv[1:6])
dt = {"bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[6] if len(sys.argv) > 6 else "bf16"]
what = sys.argv[7] if len(sys.argv) > 7 else "fwd"
m = R.RecConv2d(C, level=L).cuda()
ws = [w.detach() for w in m._param_lists()[0]]
x = torch.randn(B, C, H, W, device="cuda").to(dt); gy = torch.randn_like(x)
for _ in range(2):
    if what == "fwd": R.recconv_forward(x, ws, None, 5, L, "bilinear")
    else: R.recconv_backward(x, gy, ws, None, 5, L, "bilinear")
buf = (ctypes.c_longlong * 4096)()
lib = N.lib(); lib.recnext_debug_prof.restype = ctypes.c_int
assert lib.recnext_debug_prof(buf, 4096)
t = np.array(buf[:], dtype=np.int64); n = int((t != 0).sum()); t = t[:n]
d = np.diff(t)
print(R.plan_describe((B, C, H, W), 5, L, "bilinear", dt, False, what == "bwd"))
print("stamps", n, "total cycles", int(t[-1] - t[0]))
# stages per batch: find period by autocorrelation of the first-difference pattern over candidate periods
nstage = {"fwd": 1 + L + 2 * L + 1, "bwd": None}[what]
if what == "bwd": nstage = 1 + L + 2 * L + 1 + 2 + (1 if L > 0 else 0) + 2 * L + L
first = 1  # stamp 0 = start, first stage = filter load
per = d[first:first + (len(d) - first) // nstage * nstage].reshape(-1, nstage)
print("batches", per.shape[0], "cycles/batch median", int(np.median(per.sum(1))))
names_f = ["unpack"] + [f"down{l}" for l in range(1, L + 1)] + sum([[f"conv@{l}", f"up{l}"] for l in range(L, 0, -1)], []) + ["final"]
names_b = ["unpack x"] + [f"down{l}" for l in range(1, L + 1)] + sum([[f"conv@{l}", f"up{l}"] for l in range(L, 0, -1)], []) + \
          ["unpack gy", "wgrad final", "dgrad final"] + (["unpack x again"] if L > 0 else []) + \
          sum([[f"gather{l}", f"wgrad+dgrad@{l}"] for l in range(1, L + 1)], []) + [f"down_bwd{l}" for l in range(L, 0, -1)]
names = names_f if what == "fwd" else names_b
med = np.median(per, axis=0)
for nm, v in zip(names, med): print(f"  {nm:18s} {int(v):8d} cycles  {100 * v / med.sum():5.1f}%")
