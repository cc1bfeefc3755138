#!/usr/bin/env python
"""Per-kernel SASS instruction counts of libnnop_b200.so (cuobjdump -sass): the Blackwell
mnemonics that prove the tcgen05 / TMEM / TMA path (UTCHMMA, LDTM, STTM, UTMALDG, UTMASTG,
UTMAREDG, UTCBAR, SYNCS) next to the SIMT ones.  Runs without a GPU.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "nnop.jl_b200" / "lib" / "libnnop_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "SYNCS",
        "MUFU.EX2", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "SHFL", "STL", "LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    name, counts, order = None, {}, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(CUtensorMap_st.*", "(...)", name)
            name = re.sub(r"nnop::\(anonymous namespace\)::|void ", "", name)
            name = re.sub(r"\((?!\.\.\.).*$", "(...)", name) if len(name) > 110 else name
            counts[name] = collections.Counter()
            order.append(name)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            counts[name]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    counts[name][k] += 1
    tot = collections.Counter()
    print(f"# SASS instruction counts per kernel, {LIB.name} ({len(order)} kernels); columns with no hits omitted per row")
    for n in sorted(order, key=lambda n: -counts[n]["UTCHMMA"] * 10**6 - counts[n]["total"]):
        c = counts[n]
        tot.update(c)
        cols = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{n}\n    total={c['total']} {cols}")
    print("\n# library totals\n" + " ".join(f"{k}={tot[k]}" for k in ["total"] + KEYS if tot[k]))


if __name__ == "__main__":
    sys.exit(main())
