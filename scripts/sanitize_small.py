"""Small-shape pass over every C-ABI entry point for compute-sanitizer (development aid)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
from nnop_b200.ring import CudaBackend
torch.manual_seed(0)
dt = torch.bfloat16
for (B, QH, KH, QL, KL, E, causal) in [(1, 2, 1, 300, 300, 128, True), (1, 2, 2, 130, 257, 64, False)]:
    q = torch.randn(B, QH, QL, E, device="cuda", dtype=dt); dO = torch.randn_like(q)
    k = torch.randn(B, KH, KL, E, device="cuda", dtype=dt); v = torch.randn_like(k)
    m = torch.rand(B, KL, device="cuda") > 0.3; m[:, 0] = True
    for mask in (None, m):
        o, lse = nn._flash_attention(q, k, v, causal=causal, kpad_mask=mask)
        nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal, kpad_mask=mask)
    if E == 128:
        nn.set_bwd_pair_mode(1); nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal, kpad_mask=mask); nn.set_bwd_pair_mode(0)
# f32 generic path with pair + mask
q = torch.randn(1, 2, 70, 32, device="cuda"); k = torch.randn(1, 2, 90, 32, device="cuda"); v = torch.randn_like(k)
pr = torch.randn(1, 90, 70, 2, device="cuda"); mk = torch.ones(1, 90, dtype=torch.bool, device="cuda"); mk[0, -5:] = False
o, lse = nn._flash_attention(q, k, v, pr, causal=False, kpad_mask=mk)
nn.grad_flash_attention(torch.randn_like(q), o, lse, q, k, v, pr, causal=False, kpad_mask=mk)
# varlen
lens = [130, 1, 255, 64]
cu = torch.tensor([0, 130, 131, 386, 450], dtype=torch.int32, device="cuda")
q = torch.randn(2, 450, 128, device="cuda", dtype=dt); k = torch.randn(1, 450, 128, device="cuda", dtype=dt); v = torch.randn_like(k)
o, lse = nn._flash_attention_varlen(q, k, v, cu, cu, 255, 255, causal=True)
nn.grad_flash_attention_varlen(torch.randn_like(q), o, lse, q, k, v, cu, cu, 255, 255, causal=True)
# row-wise ops, rope, ring helpers
x = torch.randn(17, 513, device="cuda", dtype=dt); w = torch.rand(513, device="cuda", dtype=dt); b = torch.rand(513, device="cuda", dtype=dt)
y, r = nn._rms_norm(x, w); nn.grad_rms_norm(torch.randn_like(x), r, x, w)
y, mu, rs = nn._layer_norm(x, w, b); nn.grad_layer_norm(torch.randn_like(x), mu, rs, x, w, b)
y = nn.online_softmax(x); nn.grad_online_softmax(torch.randn_like(x), y)
pos = torch.arange(33, dtype=torch.float32).view(1, 33)
cos, sin = nn.LlamaRotaryEmbedding(64)(pos)
nn.llama_rope(torch.randn(1, 3, 33, 64, device="cuda", dtype=dt), torch.randn(1, 1, 33, 64, device="cuda", dtype=dt), cos=cos.cuda(), sin=sin.cuda())
be = CudaBackend()
oa = torch.zeros(1, 2, 40, 64, device="cuda"); la = torch.zeros(1, 2, 40, device="cuda")
op = torch.randn(1, 2, 40, 64, device="cuda", dtype=dt); lp = torch.randn(1, 2, 40, device="cuda")
la = be.merge(oa, la, op, lp, True); la = be.merge(oa, la, op, lp, False)
be.accumulate(oa, op, False); out = torch.empty(1, 2, 80, 64, device="cuda", dtype=dt); be.store_rows(out, oa, 40)
torch.cuda.synchronize()
print("sanitize pass done")
