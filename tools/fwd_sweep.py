"""RecConv forward: team-shape sweep on one stage shape (RECNEXT_MG / RECNEXT_MTW / RECNEXT_MNT / RECNEXT_PATH are read per call).
python tools/fwd_sweep.py B C H L"""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R
B, C, H, L = (int(v) for v in sys.argv[1:5])
m = R.RecConv2d(C, level=L).cuda()
ws = [w.detach() for w in m._param_lists()[0]]
x = torch.randn(B, C, H, H, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); R.recconv_forward(x, ws, None, 5, L, "bilinear"); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
def run(env):
    for k in ("RECNEXT_MG", "RECNEXT_MTW", "RECNEXT_MNT", "RECNEXT_PATH"): os.environ.pop(k, None)
    os.environ.update(env)
    try:
        R.recconv_forward(x, ws, None, 5, L, "bilinear"); torch.cuda.synchronize()
        t = timeit()
        d = R.plan_describe((B, C, H, H), 5, L, "bilinear", torch.bfloat16, False, False)
        print(f"{env}: {t:.4f} ms   {d[:150]}")
    except Exception as e:
        print(f"{env}: {str(e)[:120]}")
run({})
for g, tw in itertools.product((1, 2, 4, 8), (1, 2, 4)):
    run({"RECNEXT_PATH": "mma", "RECNEXT_MG": str(g), "RECNEXT_MTW": str(tw)})
