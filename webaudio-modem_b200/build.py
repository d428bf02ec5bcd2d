"""Builds libwam.so in-tree with nvcc for sm_100a (B200).  No torch involved.

The library is rebuilt whenever it was not produced from the sources in the tree: a sha256 over the CUDA sources,
the public header and the nvcc flags is stored next to the library (libwam.so.srchash) and compared, so a stale or
foreign .so (older checkout, other flags, copied file) never passes for a build of these sources."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwam.so")
STAMP = LIB + ".srchash"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    out = [os.path.join(HERE, "..", "include", "wam.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".inl", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in _sources():
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if os.environ.get("WAM_LIB"):
        return False
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        with open(STAMP) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "wam_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
