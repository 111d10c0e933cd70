"""A/B of the three exact-2x upsample + add implementations of the tensor-core forward (default: one source column per lane,
RECNEXT_MDBG=4: two source columns per lane, =8: tensor-core horizontal pass): outputs must be bit-identical."""
import os, subprocess, sys, torch
sys.path.insert(0, ".")
code = r'''
import torch, sys
sys.path.insert(0, ".")
import recnext_b200 as R
outs = []
for (B,C,H,W,L) in [(8,64,56,56,4),(8,128,28,28,3),(8,256,14,14,2),(2,16,100,168,3),(1,8,200,336,4),(2,8,96,96,4),(2,6,40,24,2),(2,6,28,14,2),(3,4,12,20,1)]:
    torch.manual_seed(B+C+H)
    x = torch.randn(B,C,H,W,device="cuda").bfloat16()
    ws = [torch.empty(C,1,5,5,device="cuda").uniform_(-0.2,0.2) for _ in range(L+2)]
    outs.append(R.recconv_forward(x, ws, None, 5, L, "bilinear").cpu())
torch.save(outs, sys.argv[1])
'''
for tag, env in (("mma", {"RECNEXT_MDBG": "8"}), ("scalar", {}), ("default", {"RECNEXT_MDBG": "4"})):
    subprocess.run([sys.executable, "-c", code, f"/tmp/ab_{tag}.pt"], check=True, env=dict(os.environ, **env))
b = torch.load("/tmp/ab_scalar.pt")
for tag in ("mma", "default"):  # "default" = the two-column variant here
    a = torch.load(f"/tmp/ab_{tag}.pt")
    for i, (u, v) in enumerate(zip(a, b)):
        print(tag, i, tuple(u.shape), "bit-identical" if torch.equal(u, v) else f"DIFF max {float((u.float()-v.float()).abs().max())} frac {float((u!=v).float().mean())}")
