"""Build ``librdetr_ops.so`` in-tree with nvcc for sm_100a (no torch headers, a few seconds).

``python -m relation_detr_b200.build`` or ``__graft_entry__.build()``.  The library is written for
B200 only: one ``-gencode arch=compute_100a,code=sm_100a``, no other architectures, no fallback.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "librdetr_ops.so")
SOURCES = ["abi.cu", "msda_fwd.cu", "msda_fwd_tile.cu", "msda_bwd.cu", "msda_bwd_coarse.cu", "msda_bwd_tile.cu", "rel.cu", "rel_attn.cu", "memfuse.cu", "topk.cu", "lsap.cu", "match_cost.cu", "diag.cu"]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC=...)")
    return cand


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "rdetr_ops.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
        "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-I", INCLUDE, "-I", CSRC,
        "-o", LIB_PATH,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    # the image's CC/CXX wrappers lack some spec files; nvcc's default host compiler (/usr/bin/g++) works
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
