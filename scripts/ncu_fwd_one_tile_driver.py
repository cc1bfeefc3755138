"""A few launches of the one-q-tile-per-CTA forward (bf16 E = 256, Float32 E = 128) for `ncu --set full`
(not a timing script): ncu ... -k regex:attn_fwd_sm100_kernel python scripts/ncu_fwd_one_tile_driver.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "nnop.jl_b200"))
import torch
import nnop_b200 as nn
q, k, v = (torch.randn(4, 16, 8192, 256, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(3):
    nn._flash_attention(q, k, v, causal=True)
assert nn.last_attention_path() == 1
q, k, v = (torch.randn(4, 16, 4096, 128, device="cuda") for _ in range(3))
for _ in range(3):
    nn._flash_attention(q, k, v, causal=True)
assert nn.last_attention_path() == 1
torch.cuda.synchronize()
print("done")
