"""Summarise an .ncu-rep (read on the CPU box): per-kernel headline metrics + opcode mix + stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__waves_per_multiprocessor"]
print(f"# ncu summary of {rep}")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"\n== {name}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:75s} {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        blocks.append([r[1], None, []])
    elif r and r[0] == "Address" and blocks:
        blocks[-1][1] = r
    elif blocks and blocks[-1][1] is not None and len(r) == len(blocks[-1][1]):
        blocks[-1][2].append(r)
for kern, h, data in blocks:
    ix = {c: i for i, c in enumerate(h)}
    ops, samp = collections.Counter(), collections.Counter()
    stalls = collections.Counter()
    scols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = tots = 0
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        op = m.group(2) if m else "?"
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "BAR", "SHFL", "UBLKCP", "SYNCS")) else op.split(".")[0]
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        ops[op] += n; samp[op] += s; tot += n; tots += s
        for c in scols:
            v = r[ix[c]]
            if v and v != "-":
                stalls[c] += int(v)
    print(f"\n== opcode mix: {kern[:100]}\n  warp instructions executed: {tot}")
    for op, n in ops.most_common(16):
        print(f"  {op:12s} {n:12d} {100 * n / tot:5.1f}%   stall samples {100 * samp[op] / max(tots, 1):5.1f}%")
    print("  stall reasons (samples):", ", ".join(f"{k[6:]}={v}" for k, v in stalls.most_common(8)))
