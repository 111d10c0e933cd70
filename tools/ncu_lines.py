"""Per-source-line totals (instructions executed, stall samples) from an .ncu-rep source page.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top N]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur_file, cur_fn, hdr = None, None, None
done = set()
agg = {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r[0] == "Line No":
        hdr = {c: i for i, c in enumerate(r)}
        iS, iI = r.index("# Samples"), r.index("Instructions Executed")
        continue
    if hdr is None or want not in (cur_fn or ""):
        continue
    if r[0] != "":  # a source line row (aggregated over its SASS)
        key = (cur_fn, cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += int(r[iI]) if r[iI].isdigit() else 0; a[1] += int(r[iS]) if r[iS].isdigit() else 0
fns = sorted({k[0] for k in agg})
for fn in fns:
    rows = [(k, v) for k, v in agg.items() if k[0] == fn]
    ti = sum(v[0] for _, v in rows); ts = sum(v[1] for _, v in rows)
    print(f"\n== {fn[:120]}\n   total instr {ti}  samples {ts}")
    for k, v in sorted(rows, key=lambda kv: -kv[1][1])[:top]:
        print(f"  {k[1]:22s}:{k[2]:4d} instr {100*v[0]/max(ti,1):5.1f}%  samples {100*v[1]/max(ts,1):5.1f}%  {v[2]}")
