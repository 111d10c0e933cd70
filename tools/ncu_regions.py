"""Samples / instructions / shared wavefronts per named source region.  python tools/ncu_regions.py rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
WANT = sys.argv[2] if len(sys.argv) > 2 else ""
fn = ""
import os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def regions_of(fname):
    """(file, first line, last line, function name) for every RC_HD / __device__ function of a csrc header"""
    path = os.path.join(ROOT, "recnext_b200", "csrc", fname)
    lines = open(path).read().split("\n")
    starts = []
    for i, ln in enumerate(lines, 1):
        m = re.match(r"^(?:RC_HD|RC_H|__device__ __forceinline__|static)\s+[\w:<>\s\*&]*?\b(\w+)\s*\(", ln)
        if m and not ln.startswith("    "):
            j = i
            while j > 1 and lines[j - 2].startswith("template"):
                j -= 1
            starts.append((j, m.group(1)))
    out = []
    for k, (ln, nm) in enumerate(starts):
        end = starts[k + 1][0] - 1 if k + 1 < len(starts) else len(lines)
        out.append((fname, ln, end, nm))
    return out


REG = []
for f in ("wstages.cuh", "recconv_stages.cuh", "wbody.cuh", "wdevice.cuh", "recconv_device.cuh"):
    REG += regions_of(f)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
hdr = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": f = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = {c: i for i, c in enumerate(r)}; continue
    if hdr is None or r[0] == "" or (WANT and WANT not in fn): continue
    g = lambda c: int(r[hdr[c]]) if (c in hdr and r[hdr[c]].isdigit()) else 0
    ln = int(r[0]); name = f
    for (ff, a, b, nm) in REG:
        if ff == f and a <= ln <= b: name = nm
    a = agg.setdefault(name, [0, 0, 0, 0])
    a[0] += g("# Samples"); a[1] += g("Instructions Executed"); a[2] += g("L1 Wavefronts Shared"); a[3] += g("L1 Wavefronts Shared Ideal")
T = [sum(a[i] for a in agg.values()) for i in range(4)]
print(f"{'region':22s} {'samples%':>8s} {'instr%':>8s} {'smem wf%':>8s}   (totals: samples {T[0]}, instr {T[1]}, wavefronts {T[2]} (ideal {T[3]}))")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:22s} {100*a[0]/max(T[0],1):8.1f} {100*a[1]/max(T[1],1):8.1f} {100*a[2]/max(T[2],1):8.1f}   wf/ideal {a[2]/max(a[3],1):.2f}")
