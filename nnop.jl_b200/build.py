#!/usr/bin/env python
"""Build libnnop_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python nnop.jl_b200/build.py [--force] [--verbose]

Output: nnop.jl_b200/lib/libnnop_b200.so (git-ignored; travels to the GPU box with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "lib" / "libnnop_b200.so"
SOURCES = ["api.cu", "rowwise.cu", "rope.cu", "attn_generic.cu", "attn_fwd_sm100.cu",
           "attn_bwd_sm100.cu", "attn_bwd_f32_sm100.cu", "attn_pair.cu", "ring_ops.cu", "ring_attn.cu", "selftest.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
         "--expt-relaxed-constexpr"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "nnop_b200.h", HERE.parent / "include" / "nnop_b200_diag.h"]

    def compile_one(src: str):
        s = CSRC / src
        o = OBJ / (src + ".o")
        if not force and not _stale(o, [s, *headers]):
            return src, 0, ""
        r = subprocess.run([NVCC, *FLAGS, "-c", str(s), "-o", str(o)], capture_output=True, text=True)
        (OBJ / (src + ".log")).write_text(r.stdout + r.stderr)
        return src, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, rc, log in results:
        if rc != 0:
            sys.stderr.write(log)
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose and log:
            print(f"==== {src}\n{log}")
    objs = [str(OBJ / (s + ".o")) for s in SOURCES]
    if force or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-cudart", "shared", "-o", str(LIB), *objs], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
