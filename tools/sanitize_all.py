"""Small invocations of every kernel of the library, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnext_b200 as R
from recnext_b200.model import dwdown_forward, ffn_forward, RecNextStem, replace_batchnorm, stem_forward, stem_pack
from recnext_b200.recattn import linattn_forward
torch.manual_seed(0)
dev = "cuda"
for (B, C, H, W, L, mode) in [(2, 8, 56, 56, 4, "bilinear"), (2, 8, 28, 28, 3, "bilinear"), (2, 8, 14, 14, 2, "bilinear"), (1, 4, 50, 84, 2, "bilinear"),
                              (1, 2, 100, 168, 3, "bilinear"), (1, 4, 25, 21, 3, "nearest"), (1, 2, 200, 336, 4, "bilinear"), (2, 8, 7, 7, 1, "bilinear")]:
    x = torch.randn(B, C, H, W, device=dev).bfloat16()
    ws = [torch.empty(C, 1, 5, 5, device=dev).uniform_(-0.2, 0.2) for _ in range(L + 2)]
    y = R.recconv_forward(x, ws, None, 5, L, mode)
    gx, gw, _ = R.recconv_backward(x, torch.randn_like(x), ws, None, 5, L, mode)   # tensor-core / FMA / streamed backward by shape
    print("recconv", (B, C, H, W, L, mode), float(y.float().abs().mean()))
x = torch.randn(2, 8, 28, 28, device=dev).bfloat16(); w = torch.randn(8, 1, 5, 5, device=dev) * 0.1; b = torch.randn(8, device=dev) * 0.1
low = R.recattn_down_forward(x, w, b); y = R.recattn_up_forward(x, low, w, b, "nearest"); y2 = R.recattn_up_forward(x, low, w, b, "bilinear")
print("recattn", float(y.float().abs().mean()), float(y2.float().abs().mean()))
xf = torch.randn(1, 4, 25, 21, device=dev)   # fp32: FMA kernels; forced streamed path
wf = [torch.empty(4, 1, 5, 5, device=dev).uniform_(-0.2, 0.2) for _ in range(4)]
os.environ["RECNEXT_PATH"] = "stream"
R.recconv_forward(xf, wf, None, 5, 2, "bilinear"); R.recconv_backward(xf, torch.randn_like(xf), wf, None, 5, 2, "nearest")
del os.environ["RECNEXT_PATH"]
print("streamed ok")
for (B, C, H) in [(2, 64, 28), (2, 128, 14), (2, 256, 14), (1, 320, 14), (1, 48, 10), (2, 512, 7), (3, 40, 9)]:
    hid = 2 * C
    yy = torch.randn(B, C, H, H, device=dev).bfloat16(); xx = torch.randn_like(yy)
    o = ffn_forward(yy, xx, torch.randn(hid, C, device=dev).bfloat16() * 0.1, torch.randn(hid, device=dev), torch.randn(C, hid, device=dev).bfloat16() * 0.1, torch.randn(C, device=dev))
    print("ffn", (B, C, H), float(o.float().abs().mean()))
for (B, C, H, W) in [(2, 64, 64, 64), (1, 80, 33, 47), (2, 40, 8, 8)]:
    st = replace_batchnorm(RecNextStem(3, C).eval().to(dev))
    o = stem_forward(torch.randn(B, 3, H, W, device=dev).bfloat16(), *stem_pack(st.stem[0], st.stem[2], torch.bfloat16))
    print("stem", (B, C, H, W), float(o.float().abs().mean()))
for (B, C, H, W) in [(2, 8, 56, 56), (3, 16, 14, 14), (1, 4, 9, 13), (40, 3, 28, 28), (2, 4, 30, 20)]:
    for dt in (torch.float32, torch.bfloat16):
        o = dwdown_forward(torch.randn(B, C, H, W, device=dev).to(dt), torch.randn(2 * C, 1, 7, 7, device=dev) / 7, torch.randn(2 * C, device=dev))
    print("dwdown", (B, C, H, W), float(o.float().abs().mean()))
for (d, heads, n) in [(32, 2, 784), (20, 2, 100), (40, 2, 49), (8, 2, 16)]:
    for dt in (torch.float32, torch.bfloat16):
        o = linattn_forward(torch.randn(2, 2 * d * heads, 1, n, device=dev).to(dt), torch.randn(2, d * heads, 1, n, device=dev).to(dt), None, heads)
    print("linattn", (d, heads, n), float(o.float().abs().mean()))
from recnext_b200.recattn import linattn_forward_pe, linattn_forward_qk
for (d, heads, h, w) in [(32, 2, 28, 28), (20, 2, 7, 9), (8, 2, 4, 4), (40, 2, 10, 13), (32, 2, 14, 14)]:
    dim = d * heads
    for dt in (torch.float32, torch.bfloat16):
        q = torch.randn(2, dim, h, w, device=dev).to(dt); k = torch.randn_like(q); v = torch.randn_like(q)
        o = linattn_forward_qk(q, k, torch.randn(dim, device=dev), torch.randn(dim, device=dev), v, torch.randn_like(v), heads)
        o = linattn_forward_pe(q, k, torch.randn(dim, device=dev), None, v, torch.randn(dim, 1, 3, 3, device=dev), torch.randn(dim, device=dev), heads)
    print("linattn qk/pe", (d, heads, h, w), float(o.float().abs().mean()))
xa = torch.randn(2, 8, 25, 21, device=dev); wa = torch.randn(8, 1, 5, 5, device=dev) * 0.1; ba = torch.randn(8, device=dev) * 0.1
lowa = R.recattn_down_forward(xa, wa, ba); ya = R.recattn_up_forward(xa, lowa, wa, ba, "bilinear")      # fp32 RecAttn pieces (gstream.cu)
print("recattn fp32", float(ya.abs().mean()))
torch.cuda.synchronize()
print("done")
