"""Build libvaeplay_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m vae_play_b200.build [--force]

One object per .cu (only stale ones are recompiled), linked with a static cudart into
vae_play_b200/lib/libvaeplay_b200.so.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libvaeplay_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "vaeplay_b200.h")
DIGEST = os.path.join(LIBDIR, "sources.sha256")     # digest of the sources the .so was built from (travels with it)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or install the CUDA toolkit)")


def _source_files():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [HEADER]


def sources_digest() -> str:
    """sha256 over the CUDA sources + the public header + the compile flags.  Content-based, so that it survives copies
    that do not preserve mtimes (the snapshot pushed to the GPU box)."""
    h = hashlib.sha256()
    for f in _source_files():
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def is_stale() -> bool:
    """True when the library is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB) or not os.path.exists(DIGEST):
        return True
    return open(DIGEST).read().strip() != sources_digest()


def _flags():
    extra = os.environ.get("VP_EXTRA_NVCC_FLAGS", "").split()       # experiments (A/B builds); empty for the product build
    return NVCC_FLAGS + (["-DVP_DEBUG_PROBES"] if os.environ.get("VP_DEBUG_PROBES") == "1" else []) + extra


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    digest = sources_digest()
    if not force and not is_stale():
        return LIB
    flagfile = os.path.join(LIBDIR, "flags.txt")
    if not os.path.exists(flagfile) or open(flagfile).read() != " ".join(_flags()):
        force = True                     # other compile flags (debug <-> release): every object is stale
    nvcc = _nvcc()
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER]
    objs = []
    procs = []
    for src in sources:
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [nvcc, *_flags(), "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[vae_play_b200.build] nvcc failed on {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"[vae_play_b200.build] {src}:\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if procs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xlinker", "--exclude-libs=ALL"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    with open(os.path.join(LIBDIR, "flags.txt"), "w") as f:
        f.write(" ".join(_flags()))
    with open(DIGEST, "w") as f:
        f.write(digest + "\n")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
