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
# ---- round 2 additions -------------------------------------------------------------------------------
# Float32 on the tensor cores (fp16 split with per-tensor scaling), E = 64 / 32 / 16, with and without pair + mask
for E in (64, 32, 16):
    q = torch.randn(1, 2, 200, E, device="cuda"); k = torch.randn(1, 1, 333, E, device="cuda"); v = torch.randn_like(k)
    pr = torch.randn(1, 333, 200, 2, device="cuda"); mk = torch.ones(1, 333, dtype=torch.bool, device="cuda"); mk[0, -7:] = False
    for args in ((None, None), (pr, mk)):
        o, lse = nn._flash_attention(q, k, v, args[0], causal=False, kpad_mask=args[1])
        assert nn.last_attention_path() == 1
        nn.grad_flash_attention(torch.randn_like(q), o, lse, q, k, v, args[0], causal=False, kpad_mask=args[1])
# 16-bit E = 32 / 16 on the E = 64 kernels (TMA zero-fill / clip), 16-bit pair bias on the tensor path
for E in (32, 16):
    q = torch.randn(1, 2, 257, E, device="cuda", dtype=dt); k = torch.randn(1, 2, 257, E, device="cuda", dtype=dt); v = torch.randn_like(k)
    o, lse = nn._flash_attention(q, k, v, causal=True)
    assert nn.last_attention_path() == 1
    nn.grad_flash_attention(torch.randn_like(q), o, lse, q, k, v, causal=True)
q = torch.randn(1, 2, 130, 128, device="cuda", dtype=dt); k = torch.randn(1, 2, 200, 128, device="cuda", dtype=dt); v = torch.randn_like(k)
pr = torch.randn(1, 200, 130, 2, device="cuda", dtype=dt)
o, lse = nn._flash_attention(q, k, v, pr, causal=False)
nn.grad_flash_attention(torch.randn_like(q), o, lse, q, k, v, pr, causal=False)
# forward / backward kernel variants: persistent (squeezed onto 3 CTAs), QUAD forward, DUO backward
q = torch.randn(2, 2, 384, 128, device="cuda", dtype=dt); k = torch.randn(2, 1, 384, 128, device="cuda", dtype=dt); v = torch.randn_like(k)
for fm, bm in ((103, 103), (3, 4), (1, 2)):
    nn.set_fwd_mode(fm); nn.set_bwd_pair_mode(bm)
    o, lse = nn._flash_attention(q, k, v, causal=True)
    nn.grad_flash_attention(torch.randn_like(q), o, lse, q, k, v, causal=True)
nn.set_fwd_mode(0); nn.set_bwd_pair_mode(0)
# ring attention behind the C ABI, two virtual ranks on this GPU (zig-zag causal and plain)
for causal in (True, False):
    qs = [torch.randn(1, 2, 256, 64, device="cuda", dtype=dt) for _ in range(2)]
    ks = [torch.randn(1, 1, 256, 64, device="cuda", dtype=dt) for _ in range(2)]
    vs = [torch.randn_like(t) for t in ks]
    os_, lses = nn.p2p_ring_attention_forward(qs, ks, vs, causal=causal)
    nn.p2p_ring_attention_backward([torch.randn_like(t) for t in qs], os_, lses, qs, ks, vs, causal=causal)
# row-wise kernels: full-row specialisations (n a multiple of the vector tile) and ragged n; second-order softmax
for n in (512, 4096, 1000):
    x = torch.randn(9, n, device="cuda", dtype=dt); w = torch.rand(n, device="cuda", dtype=dt); b = torch.rand(n, device="cuda", dtype=dt)
    y, r = nn._rms_norm(x, w); nn.grad_rms_norm(torch.randn_like(x), r, x, w)
    y, mu, rs = nn._layer_norm(x, w, b); nn.grad_layer_norm(torch.randn_like(x), mu, rs, x, w, b)
    y = nn.online_softmax(x); nn.grad_online_softmax(torch.randn_like(x), y)
xs = torch.randn(5, 300, device="cuda", requires_grad=True)
ys = nn.online_softmax(xs)
g, = torch.autograd.grad(ys, xs, torch.randn_like(ys), create_graph=True)
g.square().sum().backward()
# rope at an attention-block shape (several rows per CTA) and its backward
pos = torch.arange(300, dtype=torch.float32).view(1, 300).repeat(2, 1)
cos, sin = nn.LlamaRotaryEmbedding(128)(pos)
qr = torch.randn(2, 4, 300, 128, device="cuda", dtype=dt); kr = torch.randn(2, 2, 300, 128, device="cuda", dtype=dt)
nn.llama_rope(qr, kr, cos=cos.cuda(), sin=sin.cuda()); nn.grad_llama_rope(qr, kr, cos=cos.cuda(), sin=sin.cuda())
torch.cuda.synchronize()
print("sanitize pass done")
