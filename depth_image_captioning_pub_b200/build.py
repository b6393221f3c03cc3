"""Build the C-ABI CUDA library (libdic.so) in-tree with nvcc for sm_100a.

    python -m depth_image_captioning_pub_b200.build [--force]

The library has no torch dependency (plain pointers and sizes in every signature) and is
loaded with ctypes by ``_lib.py``.  nvcc cross-compiles without a GPU, so this also runs in
the CPU-only build container; the built ``.so`` is git-ignored but travels with the repo
snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libdic.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    out = [os.path.join(INCLUDE, "dic.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def source_hash() -> str:
    h = hashlib.sha256()
    for s in sources():
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    """Content-hash staleness (mtimes do not survive the snapshot to the GPU box)."""
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".hash"):
        return True
    with open(LIB + ".hash") as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = f"{LIB}.{os.getpid()}.tmp"       # per process: spawned multi-GPU workers may all find the library stale
    cmd = [nvcc, *NVCC_FLAGS, "-o", tmp, os.path.join(CSRC, "dic_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)
    with open(LIB + ".hash", "w") as f:
        f.write(source_hash())
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
