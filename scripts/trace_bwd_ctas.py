"""Per-CTA timeline of the backward kernel on SM 0 (variant built with -DNNOP_BWD_TRACE; development aid)."""
import ctypes, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ["NNOP_B200_LIB"] = str(ROOT / "nnop.jl_b200" / "lib" / "variants" / "libnnop_b200_btrace.so")
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
B, H, L, E = 8, 32, 8192, 128
q, k, v, dO = (torch.randn(B, H, L, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
o, lse = nn._flash_attention(q, k, v, causal=True)
buf = (ctypes.c_longlong * (16 * 1024))()
nn.grad_flash_attention(dO, o, lse, q, k, v, causal=True); torch.cuda.synchronize()
nn.lib.nnop_debug_bwd_cta_log(buf, 1024, 1)
nn.grad_flash_attention(dO, o, lse, q, k, v, causal=True); torch.cuda.synchronize()
n = nn.lib.nnop_debug_bwd_cta_log(buf, 1024, 1)
rows = sorted([list(buf[i * 16:(i + 1) * 16]) for i in range(n)], key=lambda r: r[0])
print(f"{n} CTAs ran on SM 0")
print("  j n_it |  entry  setup  first-dV  loop-end  dkdv-seen  exit | per-step  gap-to-next")
t0 = rows[0][0]; tot_steps = 0; fixed = 0
for a, b in zip(rows, rows[1:] + [None]):
    e0, e1, e2, e3, e4, e5, j, nit = a[:8]
    ep = [a[k] - e3 for k in (4, 8, 9, 10, 11, 12, 5)]
    gap = (b[0] - e5) if b else 0
    per = (e3 - e2) / max(nit, 1)
    tot_steps += nit; fixed += (e2 - e0) + (e5 - e3) + gap
    if len(rows) < 40 or rows.index(a) < 25:
        print(f"{j:3d} {nit:4d} | {e0-t0:8d} {e1-e0:6d} {e2-e0:8d} {e3-e0:9d} {e4-e0:9d} {e5-e0:7d} | {per:8.0f} {gap:8d} | after loop-end: dkdv {ep[0]} staged {ep[1]} tma-issued {ep[2]} tma-read {ep[3]} | drain last-issue {ep[4]} drained {ep[5]} | exit {ep[6]}")
span = rows[-1][5] - rows[0][0]
print(f"SM 0 span {span} clk, {tot_steps} steps; outside first-dV..loop-end: {fixed} clk = {100*fixed/span:.1f}% ({fixed/len(rows):.0f} per CTA)")
