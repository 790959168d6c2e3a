"""In-tree build of libvlsa_b200.so (nvcc, sm_100a only) and of the C oracle (gcc; test infrastructure)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libvlsa_b200.so")
SOURCES = ["api.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; vlsa_b200 needs the CUDA toolkit to build (no prebuilt fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(ROOT, "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as fh:
                    h.update(name.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into vlsa_b200/lib/libvlsa_b200.so (skipped when sources are unchanged)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest + "\n")
    return LIB


def build_oracle(verbose: bool = False) -> str | None:
    """Compile oracle/vlsa_oracle.c (the plain-C restatement used only by tests / bench cpu_baseline)."""
    mk = os.path.join(ROOT, "oracle", "Makefile")
    if not os.path.exists(mk):
        return None
    res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{res.stdout}\n{res.stderr}")
    return os.path.join(ROOT, "oracle", "_build", "libvlsa_oracle.so")


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_oracle(verbose=True))
