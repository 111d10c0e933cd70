"""Instructions executed / stall samples of one kernel aggregated per source FUNCTION REGION of a file (regions = spans
between lines matching a marker regex), plus the top source lines.   python tools/ncu_funcs.py rep file.cuh [planes]"""
import csv, io, re, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
planes = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
path = [p for p in ("recnext_b200/csrc/" + fname, fname) if __import__("os").path.exists(p)][0]
lines = open(path).read().split("\n")
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(lines) if re.match(r"^(template|__device__|struct|inline|__global__)", l) or "// ----" in l and "---------" not in l]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
hdr = None; cur = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iS, iI = r.index("# Samples"), r.index("Instructions Executed"); continue
    if hdr is None or not r[0].isdigit(): continue
    a = agg.setdefault((cur, int(r[0])), [0, 0])
    a[0] += int(r[iI]) if r[iI].isdigit() else 0; a[1] += int(r[iS]) if r[iS].isdigit() else 0
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total instr {ti} ({ti/planes:.0f}/unit) samples {ts}")
bounds = [m[0] for m in marks] + [10**9]
for k, (ln, txt) in enumerate(marks):
    i = sum(v[0] for (f, l), v in agg.items() if f == fname and ln <= l < bounds[k + 1]); s = sum(v[1] for (f, l), v in agg.items() if f == fname and ln <= l < bounds[k + 1])
    if i: print(f"  {ln:4d} {txt:70s} instr {100*i/ti:5.1f}% ({i/planes:7.0f}/unit) samples {100*s/ts:5.1f}%")
oi = sum(v[0] for (f, l), v in agg.items() if f != fname); os_ = sum(v[1] for (f, l), v in agg.items() if f != fname)
print(f"  other files: instr {100*oi/ti:5.1f}% ({oi/planes:.0f}/unit) samples {100*os_/ts:5.1f}%")
print("top lines:")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    t = lines[l - 1].strip()[:90] if f == fname and l <= len(lines) else ""
    print(f"  {f}:{l:4d} {v[0]/planes:7.0f}/unit  samples {100*v[1]/ts:4.1f}%  {t}")
