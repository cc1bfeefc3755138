"""Pipeline timeline of one forward CTA (needs a variant built with -DNNOP_FWD_TRACE; development aid)."""
import ctypes, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
name = sys.argv[1] if len(sys.argv) > 1 else "trace"
os.environ["NNOP_B200_LIB"] = str(ROOT / "nnop.jl_b200" / "lib" / "variants" / f"libnnop_b200_{name}.so")
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
E = int(os.environ.get("TRACE_E", "128"))   # 256: the one-tile kernel (stamps 11..14 = K ready, P seen, V ready, block issued)
B, H, L = 2, 74, 8192
q, k, v = (torch.randn(B, H, L, E, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(3):
    nn._flash_attention(q, k, v, causal=True)
torch.cuda.synchronize()
n = 64 * 16
buf = (ctypes.c_longlong * n)()
rc = nn.lib.nnop_debug_fwd_trace(buf, n)
assert rc == 0, rc
rows = [list(buf[i * 16:(i + 1) * 16]) for i in range(64)]
warps = os.environ.get("TRACE_WARPS") == "1"   # library built with -DNNOP_FWD_TRACE=2
names = [f"w{w}:{x}" for w in range(4) for x in ("ld", "max", "p0", "p1")] if warps else ["0:S", "0:ld", "0:max", "0:p0", "0:p1", "1:S", "1:ld", "1:max", "1:p0", "1:p1", "M0:w", "M0:i", "M1:w", "M1:i", "M0:p0", "M0:p1"]
t0 = min(rows[20][:16:4]) if warps else rows[20][0]
print("it " + " ".join(f"{x:>7s}" for x in names))
for i in range(20, 30):
    print(f"{i:2d} " + " ".join(f"{rows[i][j] - t0:7d}" for j in range(16)))
print("clk/step:", (rows[50][0] - rows[20][0]) / 30)
