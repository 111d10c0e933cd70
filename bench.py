#!/usr/bin/env python
"""bench.py — RecNeXt-M3 inference images/sec (BASELINE.json configs[1]) with the B200-native RecConv path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one forward pass of the fused-BN eval RecNeXt-M3 over a synthetic batch of 256 images (224x224, bf16
autocast) per GPU; every RecConv2d token mixer runs the fused sm_100a kernel through the C ABI.  Inference
shards by batch with no data-path collective (replicas, "weak" scaling: 256 images per GPU).
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own PyTorch CPU path on the host cores: the UNMODIFIED
reference model code staged by oracle/make_ref.py into oracle/_ref (kind "reference"), or the restatement in oracle/torch_ref.py when
that staging is absent (kind "port").

Beside the contract keys the line carries the other BASELINE.json configs as extra objects that do not disturb `value`:
`recconv_fwd_bwd` (the second half of the metric), `train_ddp` (configs[2]: RecNeXt-M5 fwd + bwd + AdamW, DDP under torchrun),
`a_series` (configs[3]: RecNeXt-A3 inference), `detection` (configs[4]: M3 backbone fwd + bwd at 800 x 1344, 2 images per GPU),
`gpu_eager_baseline` (the reference model in PyTorch eager on the same GPU: what a user gets today) and, in `cpu_baseline`,
configs[0] (RecNeXt-M0, batch 1, fp32 CPU).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL, BATCH, RES = "recnext_m3", 256, 224
METRIC, UNIT = "RecNeXt-M3 inference images/sec (batch 256/GPU, 224x224, bf16, fused-BN eval)", "images/s"
WORKLOAD = "RecNeXt-M3 inference, batch 256 at 224x224 bf16, fused-BN eval model (BASELINE.json configs[1])"


def select_model(name: str):
    """--model recnext_a3 measures BASELINE.json configs[3] (A-series) with the same contract; the default is configs[1]."""
    global MODEL, METRIC, WORKLOAD
    if name == MODEL:
        return
    MODEL = name
    tag = name.split("_")[1].upper()
    METRIC = f"RecNeXt-{tag} inference images/sec (batch 256/GPU, 224x224, bf16, fused-BN eval)"
    cfg = "configs[3]" if tag.startswith("A") else "configs[1] shape, other width/depth"
    WORKLOAD = f"RecNeXt-{tag} inference, batch 256 at 224x224 bf16, fused-BN eval model (BASELINE.json {cfg})"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        mhz, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                mhz.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(mhz)}


def build_reference_model(variant: str, device="cpu"):
    """-> (fused-BN eval model, kind).  kind "reference": the UNMODIFIED reference code (oracle/_ref, staged by oracle/make_ref.py:
    model/recnext.py + model/recattn.py + utils.replace_batchnorm, speed_gpu.py:47-50) behind the timm stand-in; kind "port": the
    restatement oracle/torch_ref.py inside this repo's model definition (same ATen kernels) when the staging is absent."""
    import torch

    torch.manual_seed(0)
    try:
        from oracle.make_ref import available, import_reference

        if available():
            create_model, replace_bn = import_reference()
            net = create_model(variant).eval()
            replace_bn(net)
            return net.to(device), "reference"
    except Exception as ex:  # fall through to the port, and say so
        sys.stderr.write(f"bench: oracle/_ref unusable ({ex}); timing the port\n")
    from oracle.torch_ref import RefRecAttn2d, RefRecConv2d
    from recnext_b200.model import create_model, replace_batchnorm

    net = create_model(variant, token_mixer=RefRecAttn2d if "_a" in variant else RefRecConv2d).eval()
    replace_batchnorm(net)
    return net.to(device), "port"


def cpu_reference_rate(steps: int, warmup: int, batch: int, variant: str = None):
    """images/sec of the reference CPU path on `batch` images per step, all host threads."""
    import torch

    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1, which would make the reference arm look 8x slower
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    net, kind = build_reference_model(variant or MODEL)
    x = torch.randn(batch, 3, RES, RES)
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            net(x)
            dt = time.perf_counter() - t0
            if i >= warmup:
                ts.append(dt)
    total = sum(ts)
    return batch * len(ts) / total, total / len(ts) * 1e3, torch.get_num_threads(), kind


def run_reference(args, rank: int):
    if rank != 0:
        return
    batch = 64   # as close to the metric's 256 as keeps a step near one second on a 16-thread host
    warm = min(args.warmup, 2)
    rate, ms, threads, kind = cpu_reference_rate(args.steps, warm, batch)
    what = "UNMODIFIED reference code (oracle/_ref)" if kind == "reference" else "restatement oracle/torch_ref.py"
    sample = f"{batch} images/step of the same model (fp32, eval, BN folded, PyTorch CPU eager, {threads} threads; {what})"
    m0_rate, m0_ms, _, _ = cpu_reference_rate(5, 2, 1, "recnext_m0")   # BASELINE.json configs[0]
    out = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "note": "CPU sample: " + sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "configs0_m0_batch1_fp32": {"images_per_s": m0_rate, "ms_per_image": m0_ms}},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def run_train(args, rank: int, local_rank: int, world: int):
    """Training step of BASELINE.json configs[2] (RecNeXt-M5 by default via --model): forward + backward + AdamW on a synthetic
    batch, bf16 autocast, one process per GPU, DDP gradient all-reduce over NCCL (the only collective of the path).
    RecConv2d runs the sm_100a kernels in both directions (tensor-core forward, fused FMA backward); the rest is PyTorch."""
    import torch
    import torch.nn.functional as F

    from recnext_b200 import dist as D
    from recnext_b200 import recconv as RC
    from recnext_b200.model import create_model

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    D.init("nccl", dev)
    torch.backends.cudnn.benchmark = True
    batch = args.batch or 128
    torch.manual_seed(0)                      # identical initial weights on every rank (DDP also broadcasts them)
    net = create_model(MODEL, drop_path=0.0).to(dev).train()
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank]) if world > 1 else net
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)
    torch.manual_seed(1 + rank)
    x = torch.randn(batch, 3, RES, RES, device=dev)
    tgt = torch.randint(0, 1000, (batch,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(model(x).float(), tgt)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    D.barrier(dev)
    RC.timing_begin()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = step()
    b.record()
    D.barrier(dev)
    ms = D.max_over_ranks(a.elapsed_time(b), dev)
    fwd = RC.timing_end()
    loss_v = float(loss.detach())
    if rank == 0:
        tag = MODEL.split("_")[1].upper()
        print(json.dumps({
            "metric": f"RecNeXt-{tag} training images/sec (fwd+bwd+AdamW, batch {batch}/GPU, 224x224, bf16 autocast)", "value": D.job_throughput(batch, world, args.steps, ms),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"RecNeXt-{tag} training step, batch {batch}/GPU at 224x224 bf16 autocast, AdamW (BASELINE.json configs[2])", "model": MODEL,
                       "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"ddp x{world} (NCCL gradient all-reduce)" if world > 1 else "single GPU"},
            "gpu_launches": len(fwd), "final_loss": loss_v,
            "recconv_forward_share_of_step": round(sum(r["ms"] for r in fwd) / ms, 4),
        }), flush=True)
    D.finalize()


def recconv_microbench(torch, R, peak):
    """RecConv fwd+bwd achieved GB/s on the four M3 stage shapes (batch 256, bf16): the second half of
    BASELINE.json's metric.  Algorithmic bytes: fwd 2*N*e, bwd 3*N*e.  L2 flushed per launch."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res, tot_bytes, tot_ms = [], 0.0, 0.0
    for shape, L in [((256, 64, 56, 56), 4), ((256, 128, 28, 28), 3), ((256, 256, 14, 14), 2), ((256, 512, 7, 7), 1)]:
        m = R.RecConv2d(shape[1], level=L).cuda()
        ws = [w.detach() for w in m._param_lists()[0]]
        x = torch.randn(shape, device="cuda").bfloat16()
        gy = torch.randn(shape, device="cuda").bfloat16()
        tf, tb = [], []
        for i in range(6):
            torch.cuda._sleep(400000)  # the GPU stays busy while the host enqueues: events bracket device time only
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); R.recconv_forward(x, ws, None, 5, L, "bilinear")
            e[1].record(); R.recconv_backward(x, gy, ws, None, 5, L, "bilinear")
            e[2].record(); torch.cuda.synchronize()
            if i >= 2:
                tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
        f, b = statistics.median(tf), statistics.median(tb)
        n = x.numel() * 2
        res.append({"shape": list(shape), "level": L, "fwd_ms": round(f, 4), "bwd_ms": round(b, 4),
                    "fwd_gbs": round(2 * n / f * 1e-6, 1), "bwd_gbs": round(3 * n / b * 1e-6, 1),
                    "fwd_bwd_gbs": round(5 * n / (f + b) * 1e-6, 1)})
        tot_bytes += 5 * n; tot_ms += f + b
    agg = tot_bytes / tot_ms * 1e-6
    return {"achieved_gbs": round(agg, 1), "frac_of_hbm_peak": round(agg / peak, 4), "per_stage": res}


def _time_steps(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run_extras(args, rank, world, dev, peak):
    """The other BASELINE.json configs, measured after the metric (extra objects of the JSON line; none of them touches `value`).
    train_ddp runs on every rank (its gradient all-reduce is a collective); the rest on rank 0 only."""
    import torch
    import torch.nn.functional as F

    from recnext_b200 import dist as D
    from recnext_b200.model import create_model, replace_batchnorm

    out = {}
    # ---- configs[2]: RecNeXt-M5 training step (fwd + bwd + AdamW, bf16 autocast), 128 images per GPU, DDP under torchrun
    try:
        torch.manual_seed(0)
        net = create_model("recnext_m5", drop_path=0.0).to(dev).train()
        ddp = world > 1
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[dev.index], gradient_as_bucket_view=True, static_graph=True) if ddp else net
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05, fused=True)
        tb = 128
        x = torch.randn(tb, 3, RES, RES, device=dev)
        tgt = torch.randint(0, 1000, (tb,), device=dev)

        def train_step(sync=True):
            opt.zero_grad(set_to_none=True)
            ctx = model.no_sync() if (ddp and not sync) else _null()
            with ctx:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = F.cross_entropy(model(x).float(), tgt)
                loss.backward()
            opt.step()

        D.barrier(dev)
        ms = D.max_over_ranks(_time_steps(torch, train_step, 4, 2), dev)
        ms_local = D.max_over_ranks(_time_steps(torch, lambda: train_step(False), 3, 1), dev) if ddp else ms
        grad_mb = sum(p.numel() for p in net.parameters()) * 4 / 1e6
        out["train_ddp"] = {"config": "BASELINE.json configs[2]: RecNeXt-M5 training step (fwd + bwd + fused AdamW), bf16 autocast, 128 images per GPU",
                            "images_per_s": tb * world / (ms * 1e-3), "ms_per_step": ms, "n_gpus": world, "global_batch": tb * world,
                            "ms_per_step_without_allreduce": ms_local, "exposed_allreduce_ms": max(ms - ms_local, 0.0), "gradient_mbytes_fp32": round(grad_mb, 1),
                            "collective": "DDP bucketed gradient all-reduce over NCCL (gradient_as_bucket_view, static_graph), overlapped with backward" if ddp else "none (1 GPU)",
                            "kernels": "RecConv2d: sm_100a tensor-core forward + backward (16-bit); BatchNorm / 1x1 convs / optimizer: PyTorch in training mode"}
        del model, net, opt, x
        torch.cuda.empty_cache()
    except Exception as ex:
        out["train_ddp"] = {"error": str(ex)[:300]}
    D.barrier(dev)
    if rank != 0:
        return out
    # ---- configs[3]: RecNeXt-A3 inference (linear attention + nearest interpolation), batch 256
    try:
        torch.manual_seed(0)
        a3 = create_model("recnext_a3").eval()
        replace_batchnorm(a3)
        a3.to(dev)
        xa = torch.randn(BATCH, 3, RES, RES, device=dev).bfloat16()

        def a3_step():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                a3(xa)

        ms = _time_steps(torch, a3_step, 8, 3)
        out["a_series"] = {"config": "BASELINE.json configs[3]: RecNeXt-A3 inference, batch 256 at 224x224 bf16, fused-BN eval model", "images_per_s": BATCH / (ms * 1e-3),
                           "ms_per_step": ms, "n_gpus": 1}
        del a3, xa
    except Exception as ex:
        out["a_series"] = {"error": str(ex)[:300]}
    # ---- configs[4]: RecNeXt-M3 backbone forward + backward at detection scale (800 x 1333 padded to 800 x 1344), 2 images per GPU, frozen BatchNorm
    try:
        torch.manual_seed(0)
        det = create_model("recnext_m3").eval().to(dev)   # eval(): BatchNorm frozen as in detection fine-tuning; gradients flow to every conv
        xd = torch.randn(2, 3, 800, 1344, device=dev, requires_grad=True)

        def det_step():
            for p_ in det.parameters():
                p_.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                f = det.forward_features(xd)
            f.float().mean().backward()

        ms = _time_steps(torch, det_step, 3, 2)
        out["detection"] = {"config": "BASELINE.json configs[4]: RecNeXt-M3 backbone fwd + bwd at 800x1344 (RecConv level 4 at stage 0), 2 images per GPU, bf16 autocast",
                            "images_per_s": 2 / (ms * 1e-3), "ms_per_step": ms, "n_gpus": 1,
                            "recconv_paths": "stage 0 (200x336): tensor-core forward, streamed backward; stage 1 (100x168) and below: tensor-core forward and backward on chip"}
        del det, xd
        torch.cuda.empty_cache()
    except Exception as ex:
        out["detection"] = {"error": str(ex)[:300]}
    # ---- what a user of the reference gets on this GPU today: the reference model in PyTorch eager (cuDNN / ATen), same config as the metric
    try:
        ref, kind = build_reference_model(MODEL, dev)
        xr = torch.randn(BATCH, 3, RES, RES, device=dev).bfloat16()
        torch.backends.cudnn.benchmark = True

        def ref_step():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                ref(xr)

        ms = _time_steps(torch, ref_step, 5, 3)
        out["gpu_eager_baseline"] = {"config": WORKLOAD + " — PyTorch eager on the same GPU", "images_per_s": BATCH / (ms * 1e-3), "ms_per_step": ms,
                                     "kind": kind, "note": "cudnn.benchmark=True as in the reference harness (main.py:213); 4L+2 launches per RecConv2d, every intermediate through HBM"}
        del ref, xr
        torch.cuda.empty_cache()
    except Exception as ex:
        out["gpu_eager_baseline"] = {"error": str(ex)[:300]}
    return out


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra objects (train_ddp, a_series, detection, gpu_eager_baseline)")
    ap.add_argument("--no-graph", action="store_true", help="e2e leg: launch the kernels eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--model", default=MODEL, help="recnext_m0..m5 / recnext_a0..a5 (default: the metric's recnext_m3)")
    ap.add_argument("--train", action="store_true",
                    help="BASELINE.json configs[2]: training step (fwd + bwd + AdamW, bf16 autocast, DDP over NCCL when launched with torchrun) "
                         "instead of inference; --batch images per GPU (default 128)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--memory-format", default="contiguous", choices=["contiguous", "channels_last"],
                    help="memory format of the model around RecConv2d (the RecConv kernels always work on NCHW planes)")
    args = ap.parse_args()
    select_model(args.model)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.train:
        run_train(args, rank, local_rank, world)
        return

    import torch

    import recnext_b200 as R
    from recnext_b200 import dist as D
    from recnext_b200 import recconv as RC
    from recnext_b200.model import create_model, replace_batchnorm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the RecConv path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    D.init("nccl", dev)
    torch.backends.cudnn.benchmark = True  # as the reference harness does (main.py:213)

    torch.manual_seed(0 + rank)
    net = create_model(MODEL).eval()
    replace_batchnorm(net)  # the "fused-BN eval model" (speed_gpu.py:48)
    net.to(dev)
    cl = args.memory_format == "channels_last"
    if cl:
        net.to(memory_format=torch.channels_last)
    x_dev = torch.randn(BATCH, 3, RES, RES, device=dev).bfloat16()
    if cl:
        x_dev = x_dev.contiguous(memory_format=torch.channels_last)
    x_host = torch.randn(BATCH, 3, RES, RES).bfloat16().pin_memory()
    y_host = torch.empty(BATCH, 1000, dtype=torch.bfloat16).pin_memory()
    x_stage = torch.empty_like(x_dev)

    def step_device():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return net(x_dev)

    def step_e2e():
        x_stage.copy_(x_host, non_blocking=True)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            y = net(x_stage.contiguous(memory_format=torch.channels_last) if cl else x_stage)
        y_host.copy_(y, non_blocking=True)

    def barrier():
        D.barrier(dev)

    def timed(fn, steps, instrument=False):
        barrier()
        if instrument:
            RC.timing_begin()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        launches = RC.timing_end() if instrument else None
        ms = D.max_over_ranks(ms, dev)
        return ms, launches

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, _ = timed(step_device, args.steps)                    # the metric: no per-launch instrumentation in the timed region
    n_inst = min(args.steps, 5)
    ms_inst, launches = timed(step_device, n_inst, instrument=True)  # second pass: CUDA events around every launch of this repo's kernels
    # e2e: the repo's host-side inference loop (recnext_b200.infer.PipelinedInference): every step copies its batch from pinned
    # host memory and its logits back; the copy of batch i+1 overlaps the compute of batch i (separate copy stream)
    from recnext_b200.infer import PipelinedInference

    runner = PipelinedInference(net, torch.bfloat16, dev, cuda_graph=not args.no_graph)

    def run_e2e(steps):
        n = 0
        for _y in runner.run(x_host for _ in range(steps)):
            n += 1
        return n

    if cl:
        for _ in range(2):
            step_e2e()
        ms_e2e, _ = timed(step_e2e, args.steps)
    else:
        run_e2e(2)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        assert run_e2e(args.steps) == args.steps
        eb.record()
        barrier()
        ms_e2e = D.max_over_ranks(ea.elapsed_time(eb), dev)
    clocks = sampler.stop() if rank == 0 else None

    value = D.job_throughput(BATCH, world, args.steps, ms_total)
    e2e_value = D.job_throughput(BATCH, world, args.steps, ms_e2e)

    # roofline of the dominant kernel of the step: the tensor-core fused RecConv forward (19 of the 21 RecConv launches
    # of an M3 step; the two 7x7 stage-3 launches take the FMA kernel and are listed beside it)
    peak, peak_src = hbm_peak()
    groups = {}
    for rec in launches:
        g = groups.setdefault(rec["shape"], {"bytes": 0, "ms": 0.0, "n": 0})
        g["bytes"] += rec["bytes"]; g["ms"] += rec["ms"]; g["n"] += 1
    per_shape, dom = [], {"bytes": 0, "ms": 0.0, "n": 0}
    for shape, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
        if isinstance(shape[0], str):   # RecAttn2d pieces: ("down" | "up", B, C, H, W)
            kern = {"ffn": "recnext_ffn_tc_kernel", "dwdown": "recnext_dwdown16_kernel", "stem": "recnext_stem_kernel"}.get(shape[0], "recconv_mfwd_kernel/" + shape[0])
        else:
            level = {56: 4, 28: 3, 14: 2, 7: 1}.get(shape[2], 0) if RES == 224 else None
            desc = RC.plan_describe(shape, 5, level, "bilinear", torch.bfloat16, False, False) if level is not None else ""
            kern = "recconv_mfwd_static_kernel" if "compile-time" in desc else ("recconv_mfwd_kernel" if "tensor-core" in desc else "recconv_wfwd_kernel")
        gbs = g["bytes"] / (g["ms"] * 1e-3) * 1e-9 if g["ms"] > 0 else 0.0
        per_shape.append({"shape": list(shape), "kernel": kern, "launches_per_step": g["n"] / max(n_inst, 1),
                          "avg_launch_ms": round(g["ms"] / g["n"], 5), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)})
        if kern.startswith("recconv_mfwd") or ("_a" in MODEL and kern.startswith("recconv_mfwd_kernel/")):
            dom["bytes"] += g["bytes"]; dom["ms"] += g["ms"]; dom["n"] += g["n"]
    if dom["n"] == 0:
        dom = {"bytes": sum(r["bytes"] for r in launches), "ms": sum(r["ms"] for r in launches), "n": len(launches)}
    kern_all_ms = sum(rec["ms"] for rec in launches)
    n_launch = int(round(len(launches) / max(n_inst, 1) * args.steps))   # launches of this repo's kernels in the K timed steps (counted in the instrumented pass)
    achieved = dom["bytes"] / (dom["ms"] * 1e-3) * 1e-9 if dom["ms"] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get("recconv_fwd_dram_bytes_per_launch_avg")
    except Exception:
        pass
    roofline = {
        "bound": "hbm",
        "kernel": "recnext::recconv_mfwd%s_kernel<bf16> (tensor-core fused %s forward; its %d launches of the timed steps)"
                  % (("", "RecAttn2d down / up-add-conv", dom["n"]) if "_a" in MODEL else ("_static", "RecConv", dom["n"])),
        "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
        "peak_source": peak_src, "bytes_per_launch": dom["bytes"] / max(dom["n"], 1), "avg_launch_ms": dom["ms"] / max(dom["n"], 1),
        "share_of_step": round(dom["ms"] / ms_inst, 4), "all_custom_kernels_share_of_step": round(kern_all_ms / ms_inst, 4),
        "measured_in": f"a second pass of {n_inst} steps with an event pair around every launch (the timed `value` pass carries no instrumentation)",
        "per_shape": per_shape,
        "note": "algorithmic bytes 2*N*e per launch (SURVEY 8d), CUDA events on the launching stream around every launch of the "
                "timed steps.  The block is not HBM bound on B200: ~48 MAC per 4 bytes of bf16 traffic; the stencils run on the "
                "tensor cores as banded-Toeplitz MMAs, bound by shared-memory bandwidth and issue slots (DESIGN.md 3.2)",
    }

    # second roofline: the fused channel-mixer kernel (by time the largest kernel of the step since RecConv got fast).  It is a
    # pair of GEMMs, so it is reported against the tensor peak too: flops = 2 GEMMs x 2 x C x hidden per pixel.
    ffn = [r for r in launches if isinstance(r["shape"][0], str) and r["shape"][0] == "ffn"]
    roofline_ffn = None
    if ffn:
        hid_ratio = 2.0 if "_a" not in MODEL else 1.875
        fl = sum(4.0 * r["shape"][2] * (r["shape"][2] * hid_ratio) * r["shape"][1] * r["shape"][3] * r["shape"][4] for r in ffn)
        fms = sum(r["ms"] for r in ffn)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                tpeak, tsrc = float(json.load(fh)["bf16_tflops_sustained"]), "measured sustained (MEASURED_PEAKS.json)"
        except Exception:
            tpeak, tsrc = 1400.0, "fallback (B200_PROFILING.md)"
        tf = fl / (fms * 1e-3) * 1e-12
        gbs = sum(r["bytes"] for r in ffn) / (fms * 1e-3) * 1e-9
        roofline_ffn = {"bound": "tensor", "kernel": "recnext::recnext_ffn_tc_kernel<bf16> (fused channel mixer on tcgen05 / TMEM / TMA; %d launches)" % len(ffn),
                        "achieved": round(tf, 1), "peak": tpeak, "unit": "TFLOP/s", "frac": round(tf / tpeak, 4), "peak_source": tsrc,
                        "hbm_gbs_of_3Ne": round(gbs, 1), "hbm_frac": round(gbs / peak, 4), "avg_launch_ms": fms / len(ffn),
                        "share_of_step": round(fms / ms_inst, 4),
                        "note": "tcgen05.mma kernel (TMEM accumulators, TMA-streamed weight tiles, writer warps for the residual + stores): the narrow stages "
                                "are bound by the epilogue warps' instruction issue (GELU), the wide ones by the MMA issuer's instruction chain "
                                "(345 clocks per 4-MMA ring slot against 256 of math, DESIGN.md 3.4)"}

    extras = {} if args.no_extras else run_extras(args, rank, world, dev, peak)

    if rank != 0:
        D.finalize()
        return

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "variant": MODEL, "batch_per_gpu": BATCH, "global_batch": BATCH * world, "resolution": RES,
                   "parallelism": f"replicas x{world} (no data-path collective)", "weights": "random-init",
                   "memory_format": args.memory_format + " (NCHW planes for every kernel of this repo: stem, RecConv, channel mixers and downsample convs of all four stages; library kernels only for the pooled head)",
                   "l2": "per-step activations (>= 100 MB per stage-0 tensor) exceed the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": x_host.numel() * 2,
                "d2h_bytes_per_step": y_host.numel() * 2, "api": "recnext_b200.infer.PipelinedInference(model%s).run(pinned host batches) -> pinned host logits (H2D of batch i+1 overlaps compute of batch i)" % ("" if args.no_graph else ", cuda_graph=True")},
        "gpu_launches": n_launch,
        "roofline": roofline,
    }
    if roofline_ffn is not None:
        out["roofline_channel_mixer"] = roofline_ffn
    if not args.no_micro and "_a" not in MODEL:
        out["recconv_fwd_bwd"] = recconv_microbench(torch, R, peak)
    out.update(extras)
    if not args.no_cpu_baseline:
        rate, ms, threads, kind = cpu_reference_rate(steps=3, warmup=1, batch=32)
        m0_rate, m0_ms, _, _ = cpu_reference_rate(5, 2, 1, "recnext_m0")
        what = "UNMODIFIED reference code (oracle/_ref)" if kind == "reference" else "restatement oracle/torch_ref.py"
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": f"32 images/step x 3 steps of the same model (fp32, eval, BN folded, PyTorch CPU eager, {threads} threads; {what})",
                               "configs0_m0_batch1_fp32": {"images_per_s": m0_rate, "ms_per_image": m0_ms}}
    print(json.dumps(out), flush=True)
    D.finalize()


if __name__ == "__main__":
    main()
