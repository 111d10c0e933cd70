"""Per-stage RecConv kernel timings (CUDA events, L2 flushed between iterations) vs the HBM roofline.

    python tools/kbench.py [--dtype bf16|f32] [--eager] [--shapes m3|m5|det]

Algorithmic bytes: forward 2*N*e, backward 3*N*e.  Not the bench contract (bench.py is).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import recnext_b200 as R  # noqa: E402

SHAPES = {
    "m3": [((256, 64, 56, 56), 4), ((256, 128, 28, 28), 3), ((256, 256, 14, 14), 2), ((256, 512, 7, 7), 1)],
    "m5": [((128, 80, 56, 56), 4), ((128, 160, 28, 28), 3), ((128, 320, 14, 14), 2), ((128, 640, 7, 7), 1)],
    "det": [((2, 128, 100, 168), 3), ((2, 256, 50, 84), 2), ((2, 512, 25, 42), 1)],
}


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        torch.cuda._sleep(400000)  # keeps the GPU busy while the host enqueues: events then bracket device time only
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--shapes", default="m3")
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--mode", default="bilinear")
    args = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[args.dtype]
    e = 2 if dt != torch.float32 else 4
    pk, how = peak()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f"# HBM peak {pk} GB/s ({how}); dtype {args.dtype}; mode {args.mode}")
    for shape, L in SHAPES[args.shapes]:
        C = shape[1]
        m = R.RecConv2d(C, level=L, mode=args.mode).to(dev)
        ws = [w.detach() for w in m._param_lists()[0]]
        x = torch.randn(shape, device=dev).to(dt)
        gy = torch.randn(shape, device=dev).to(dt)
        N = x.numel()
        try:
            tf = timeit(lambda: R.recconv_forward(x, ws, None, 5, L, args.mode), flush)
            tb = timeit(lambda: R.recconv_backward(x, gy, ws, None, 5, L, args.mode), flush)
        except RuntimeError as ex:
            print(shape, "unsupported:", ex)
            continue
        gf, gb = 2 * N * e / tf * 1e-6, 3 * N * e / tb * 1e-6
        gfb = 5 * N * e / (tf + tb) * 1e-6
        line = (f"{str(shape):22s} L={L} fwd {tf:7.3f} ms {gf:7.0f} GB/s ({gf / pk:5.1%}) | bwd {tb:7.3f} ms {gb:7.0f} GB/s "
                f"({gb / pk:5.1%}) | fwd+bwd {gfb:7.0f} GB/s ({gfb / pk:5.1%})")
        if args.eager:
            from oracle.torch_ref import recconv_reference

            xr = x.clone().requires_grad_(True)
            wr = [w.to(dt).requires_grad_(True) for w in ws]

            def eager_f():
                with torch.no_grad():
                    recconv_reference(x, wr[0], wr[1:], None, None, args.mode)

            def eager_fb():
                y = recconv_reference(xr, wr[0], wr[1:], None, None, args.mode)
                y.backward(gy)

            te = timeit(eager_f, flush, iters=5)
            teb = timeit(eager_fb, flush, iters=5)
            line += f" | eager fwd {te:7.3f} ms fwd+bwd {teb:7.3f} ms (x{te / tf:.1f}, x{teb / (tf + tb):.1f})"
        print(line)
        print("   ", R.plan_describe(shape, 5, L, args.mode, dt, False, False))
        print("   ", R.plan_describe(shape, 5, L, args.mode, dt, False, True))


if __name__ == "__main__":
    main()
