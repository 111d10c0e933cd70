"""Samples / instructions / shared wavefronts per named source region.  python tools/ncu_regions.py rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
REG = [("wstages.cuh", 13, 27, "item decode"), ("wstages.cuh", 32, 64, "conv_s1"), ("wstages.cuh", 66, 102, "conv_s2"),
       ("wstages.cuh", 104, 116, "store_level"), ("wstages.cuh", 117, 170, "up2x"), ("wstages.cuh", 171, 215, "up generic"),
       ("wstages.cuh", 216, 235, "store_global"), ("wstages.cuh", 236, 260, "up_bwd gather"), ("wstages.cuh", 261, 300, "wgrad_s1"),
       ("wstages.cuh", 301, 345, "wgrad_s2"), ("wstages.cuh", 346, 420, "convT_s2"),
       ("recconv_stages.cuh", 84, 93, "load_row"), ("recconv_stages.cuh", 94, 131, "unpack"), ("recconv_stages.cuh", 132, 150, "load_filter"),
       ("recconv_stages.cuh", 59, 83, "elem cvt")]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
hdr = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": f = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = {c: i for i, c in enumerate(r)}; continue
    if hdr is None or r[0] == "" or r[0] == "Function Name": continue
    g = lambda c: int(r[hdr[c]]) if r[hdr[c]].isdigit() else 0
    ln = int(r[0]); name = f
    for (ff, a, b, nm) in REG:
        if ff == f and a <= ln <= b: name = nm
    a = agg.setdefault(name, [0, 0, 0, 0])
    a[0] += g("# Samples"); a[1] += g("Instructions Executed"); a[2] += g("L1 Wavefronts Shared"); a[3] += g("L1 Wavefronts Shared Ideal")
T = [sum(a[i] for a in agg.values()) for i in range(4)]
print(f"{'region':22s} {'samples%':>8s} {'instr%':>8s} {'smem wf%':>8s}   (totals: samples {T[0]}, instr {T[1]}, wavefronts {T[2]} (ideal {T[3]}))")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:22s} {100*a[0]/max(T[0],1):8.1f} {100*a[1]/max(T[1],1):8.1f} {100*a[2]/max(T[2],1):8.1f}   wf/ideal {a[2]/max(a[3],1):.2f}")
