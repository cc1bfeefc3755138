"""Per-tile timeline of the persistent backward kernel's CTA 0 (variant built with -DNNOP_BWD_TRACE;
development aid):  python scripts/build_variants.py attn_bwd_sm100.cu btrace:-DNNOP_BWD_TRACE"""
import ctypes, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("NNOP_B200_LIB", str(ROOT / "nnop.jl_b200" / "lib" / "variants" / "libnnop_b200_btrace.so"))
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
causal = "--noncausal" not in sys.argv
B, H, L, E = (8, 32, 8192, 128) if causal else (2, 32, 8192, 128)
q, k, v, dO = (torch.randn(B, H, L, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
o, lse = nn._flash_attention(q, k, v, causal=causal)
buf = (ctypes.c_longlong * (16 * 1024))()
nn.set_bwd_pair_mode(3)
for _ in range(2):
    nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal); torch.cuda.synchronize()
n = nn.lib.nnop_debug_bwd_cta_log(buf, 1024, 1)
rows = [list(buf[i * 16:(i + 1) * 16]) for i in range(n)]
print(f"CTA 0 walked {n} tiles")
print("tile  j n_it | since prev end: start dP0 S0 dV0 dQ0 dK0 dV1 | steady per-step | tile clk")
tot = steps = bnd = 0
prev_end = None
for i, r in enumerate(rows):
    st, dp0, s0, dv0, dq0, dk0, end, nit, j, dv1 = r[:10]
    base = prev_end if prev_end is not None else st
    per = (end - dv1) / max(nit - 1, 1) if nit > 1 else 0
    if 0 < i < 30:
        pe = rows[i - 1]
        # drain / compute stamps: epilogue of the PREVIOUS tile relative to its end; this tile's first P^T
        print(f"      prev tile epilogue: dQ loop done {pe[14]-base} dv_empty {pe[10]-base} dk_empty {pe[11]-base} | this tile: s_full seen {r[13]-base} p_full arrive {r[12]-base}")
    if i < 30:
        print(f"{i:4d} {j:3d} {nit:4d} | {st-base:6d} {dp0-base:6d} {s0-base:6d} {dv0-base:6d} {dq0-base:6d} {dk0-base:6d} {dv1-base:6d} | {per:7.0f} | {end-base:8d}")
    if prev_end is not None and nit > 1:
        bnd += dv1 - prev_end; steps += nit - 1; tot += end - dv1
    prev_end = end
print(f"mean steady step {tot / max(steps, 1):.0f} clk over {steps} steps; mean boundary (prev end -> second dV) {bnd / max(n - 1, 1):.0f} clk")
print(f"CTA span {rows[-1][6] - rows[0][0]} clk")
