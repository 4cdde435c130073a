"""Vendor the UNMODIFIED reference modules of the hot path into oracle/_ref/ (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

    python oracle/build_ref.py          # build container only: needs /root/reference (read-only)

The reference is pure Python, so "building" it is copying the files the path needs, byte for byte, from where they lie
under /root/reference into oracle/_ref/ -- which is git-ignored (never part of the history) but not gpurun-ignored, so it
travels to the GPU box next to the built .so files.  There ``bench.py --impl reference`` and the ``cpu_baseline`` leg time
these modules on the host cores (``kind: "reference"``); without them they fall back to the line-by-line port
oracle/vae_torch.py (``kind: "port"``).  A SHA-256 manifest records what was copied.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")
FILES = ["models/networks.py", "models/blocks.py", "models/network_Style_GAN.py", "tools/ops.py"]


def build_ref() -> bool:
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(OUT, "models", "networks.py"))
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref", "ready" if build_ref() else "unavailable (no /root/reference)")
