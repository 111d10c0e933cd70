"""Top-sampled SASS instructions with their dominant stall reasons.  python tools/ncu_sass.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {c: i for i, c in enumerate(h)}
scols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instrs", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[c]]) if r[ix[c]].isdigit() else 0, c[6:]) for c in scols), reverse=True)[:3]
    print(f"{i:6d} {r[ix['Source']][:70]:70s} smp {int(r[ix['# Samples']]):6d} exec {r[ix['Instructions Executed']]:>9s}  " + " ".join(f"{n}:{v}" for v, n in st if v))
