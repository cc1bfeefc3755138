"""Dump the per-step pipeline timeline of one backward CTA (NNOP_BWD_TRACE=<file>)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
out = str(ROOT / "gpurun_out" / "bwd_trace.txt")
os.environ["NNOP_BWD_TRACE"] = out
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
B, H, L, E = 4, 16, 4096, 128
q, k, v, dO = (torch.randn(B, H, L, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
o, lse = nn._flash_attention(q, k, v, causal=True)
for _ in range(2):
    nn.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
torch.cuda.synchronize()
names = ["A:dV", "A:S+1", "B:dQdK", "B:dP+1", "W1:S", "W1:P", "W2:P", "W2:dP", "W2:dS", "D:dQ", "D:rd"]
rows = [list(map(int, l.split())) for l in open(out)]
print("it " + " ".join(f"{n:>8s}" for n in names))
for r in rows[4:14]:
    print(f"{r[0]:2d} " + " ".join(f"{x:8d}" for x in r[1:]))
r0, r1 = rows[5], rows[13]
print("avg clk/step (A:dV):", (r1[1] - r0[1]) / 8)
