"""One forward + backward with a pair bias at the reference's benchmark shape (for an ncu launch list)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch
import nnop_b200 as nn
B, H, L, E = 4, 4, 2048, 64
dt = torch.float32 if "--bf16" not in sys.argv else torch.bfloat16
q, k, v, dO = (torch.randn(B, H, L, E, device="cuda", dtype=dt) for _ in range(4))
pair = torch.randn(B, L, L, H, device="cuda", dtype=dt)
for _ in range(2):
    o, lse = nn._flash_attention(q, k, v, pair, causal=False)
    nn.grad_flash_attention(dO, o, lse, q, k, v, pair, causal=False)
torch.cuda.synchronize()
print("path", nn.last_attention_path())
