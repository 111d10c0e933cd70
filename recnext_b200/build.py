"""Builds librecnext_b200.so (hand-written sm_100a CUDA + the C ABI of include/recnext_b200.h) in-tree.

    python -m recnext_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The library is written to recnext_b200/_lib/ (git-ignored,
but it travels to the GPU box with the working-tree snapshot).  Objects go to build/ (git-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "librecnext_b200.so")
OBJDIR = os.path.join(ROOT, "build", "obj")

SOURCES = ["capi.cu", "recconv_k3.cu", "recconv_k5.cu", "recconv_k7.cu", "recconv_w3.cu", "recconv_w5.cu", "recconv_w7.cu", "recconv_m5.cu", "recconv_mb5.cu", "ffn_mma.cu", "dwdown.cu", "linattn.cu", "linattn_mma.cu", "gstream.cu", "ffn_tc.cu", "stem.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA toolkit is required to build recnext_b200 (no CPU fallback exists)")
    return exe


def _deps():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(ROOT, "include", "recnext_b200.h"))
    files.append(os.path.abspath(__file__))
    return files


def _stamp() -> str:
    h = hashlib.sha256()
    for f in _deps():
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    cmd = [nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read().strip() == stamp:
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    cmd = [nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    try:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    except RuntimeError as ex:
        print(str(ex)[-6000:])
        print("BUILD FAILED")
        sys.exit(1)
