"""One launch of every HBM-bound kernel at BASELINE config C3's shapes (and the reference's softmax
benchmark shape), for `ncu --set full` (scripts/ncu_bandwidth.sh).  Not a timing script."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch  # noqa: E402
import nnop_b200 as nn  # noqa: E402

dt = torch.bfloat16
n, emb = 8192, 4096
x = torch.rand(n, emb, device="cuda").to(dt)
dy = torch.randn(n, emb, device="cuda").to(dt)
w = torch.rand(emb, device="cuda").to(dt)
b = torch.rand(emb, device="cuda").to(dt)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    y, rstd = nn._rms_norm(x, w)
    nn.grad_rms_norm(dy, rstd, x, w)
    y, mu, rs = nn._layer_norm(x, w, b)
    nn.grad_layer_norm(dy, mu, rs, x, w, b)
    xs = torch.randn(1024, 8192, device="cuda").to(dt)
    ys = nn._softmax_fwd(xs) if hasattr(nn, "_softmax_fwd") else nn.online_softmax(xs)
    nn.grad_online_softmax(torch.randn_like(xs), ys)
    B, QH, KH, L, E = 1, 32, 8, 8192, 128
    q = torch.randn(B, QH, L, E, device="cuda").to(dt)
    k = torch.randn(B, KH, L, E, device="cuda").to(dt)
    v = torch.randn(B, KH, L, E, device="cuda").to(dt)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = (t.cuda() for t in nn.LlamaRotaryEmbedding(E)(pos))
    qr, kr = nn.llama_rope(q, k, cos=cos, sin=sin)
    o, lse = nn._flash_attention(qr, kr, v, causal=True)
    nn.grad_flash_attention(torch.randn_like(o), o, lse, qr, kr, v, causal=True)   # prep / post kernels
    # the reference's benchmark shape with a pair bias (benchmarks/main.jl:305-386: Float32 E=64 L=2048 H=4 B=4):
    # layout kernels of the bias (pair -> head-major, dpair back), |x|max pass and fp16 split of the Float32 path
    g = torch.Generator(device="cuda").manual_seed(0)
    Bp, Hp, Lp, Ep = 4, 4, 2048, 64
    qf, kf, vf, dOf = (torch.randn(Bp, Hp, Lp, Ep, device="cuda", generator=g) for _ in range(4))
    pr = torch.randn(Bp, Lp, Lp, Hp, device="cuda", generator=g)
    of, lsef = nn._flash_attention(qf, kf, vf, pr, causal=True)
    nn.grad_flash_attention(dOf, of, lsef, qf, kf, vf, pr, causal=True)
    torch.cuda.synchronize()
print("ok")
