"""Build libhoisdf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hoisdf_b200.csrc.build [--force] [--verbose]

Objects go to hoisdf_b200/csrc/build/, the library to hoisdf_b200/libhoisdf_b200.so (git-ignored, but it
travels to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(PKG, "libhoisdf_b200.so")

SOURCES = ["api.cu", "linear.cu", "linear_tc.cu", "linear_tc2.cu", "linear_h3.cu", "resnet.cu", "narrow.cu", "lattice.cu", "gather.cu", "sdf.cu", "sdf_chain.cu", "sdf_infer.cu", "topk.cu", "attention.cu", "attention_tc.cu", "layernorm.cu", "transformer.cu", "heads.cu", "metrics.cu", "backward.cu", "train_prep.cu", "feed.cu", "augment.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--fmad=true",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libhoisdf_b200.so")


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))) + ["build.py"]
    for f in files:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    with open(os.path.join(ROOT, "include", "hoisdf_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
