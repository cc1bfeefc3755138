#!/usr/bin/env python
"""Build experiment variants of libnnop_b200.so: one source recompiled with extra -D flags, linked
against the regular objects.  python scripts/build_variants.py <src.cu> name1:-DX=1,-DY=2 name2:...
Output: nnop.jl_b200/lib/variants/libnnop_b200_<name>.so (select with NNOP_B200_LIB)."""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "nnop.jl_b200"
sys.path.insert(0, str(PKG))
import build as B
B.build()
src = sys.argv[1]
out = PKG / "lib" / "variants"; out.mkdir(parents=True, exist_ok=True)
procs = []
for spec in sys.argv[2:]:
    name, _, defs = spec.partition(":")
    o = B.OBJ / f"{src}.{name}.o"
    cmd = [B.NVCC, *B.FLAGS, *[d for d in defs.split(",") if d], "-c", str(B.CSRC / src), "-o", str(o)]
    procs.append((name, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, o, pr in procs:
    log, _ = pr.communicate()
    if pr.returncode: sys.exit(log)
    (B.OBJ / f"{src}.{name}.log").write_text(log)
    objs = [str(o) if s == src else str(B.OBJ / (s + ".o")) for s in B.SOURCES]
    lib = out / f"libnnop_b200_{name}.so"
    subprocess.run([B.NVCC, "-shared", "-cudart", "shared", "-o", str(lib), *objs], check=True)
    print(lib)
