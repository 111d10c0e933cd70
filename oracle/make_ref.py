"""Stages the UNMODIFIED reference model code for the CPU arm of bench.py (TEST / BASELINE INFRASTRUCTURE, never imported by the product).

    python oracle/make_ref.py            (run by __graft_entry__.build() wherever /root/reference exists)

The reference is a pure-Python script repository (no setup.py: `pip install` refuses it) whose hot path is
model/recnext.py + model/recattn.py, with utils.replace_batchnorm for the fused-BN eval model (speed_gpu.py:47-50).  This
recipe copies those files VERBATIM from /root/reference into oracle/_ref/refsrc/ — git-ignored, so no reference source enters
the history, but part of the working-tree snapshot that travels to the GPU box.  `timm` (third party, absent) is provided
by the 3-file stand-in oracle/timm_shim.  A manifest with SHA-256 digests records what was staged.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref", "refsrc")
FILES = ["model/__init__.py", "model/recnext.py", "model/recattn.py", "utils.py"]


def make(ref: str = REF, out: str = OUT) -> bool:
    if not os.path.isdir(ref):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(ref, rel), os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(out, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref, "files": manifest}, fh, indent=1)
    return True


def available(out: str = OUT) -> bool:
    return all(os.path.exists(os.path.join(out, rel)) for rel in FILES)


def import_reference(out: str = OUT):
    """-> (create_model, replace_batchnorm) of the staged reference; raises if oracle/_ref was not made."""
    if not available(out):
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    for p in (os.path.join(HERE, "timm_shim"), out):
        if p not in sys.path:
            sys.path.insert(0, p)
    import model  # noqa: F401  (the reference package: registers recnext_m0..m5 / recnext_a0..a5 with the shim's registry)
    import utils as ref_utils
    from timm.models import create_model

    return create_model, ref_utils.replace_batchnorm


if __name__ == "__main__":
    ok = make()
    print("staged" if ok else "no /root/reference here", OUT if ok else "")
