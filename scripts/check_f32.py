import sys, math
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200")); sys.path.insert(0, str(ROOT))
import torch, nnop_b200 as nn
from oracle import oracle as O
torch.manual_seed(0)
for (B, QH, KH, QL, KL, causal) in [(1, 2, 2, 256, 256, False), (2, 2, 1, 255, 511, False), (1, 2, 2, 1024, 1024, True), (3, 2, 2, 512, 512, True)]:
    q = torch.randn(B, QH, QL, 64); k = torch.randn(B, KH, KL, 64); v = torch.randn(B, KH, KL, 64)
    o, lse = nn._flash_attention(q.cuda(), k.cuda(), v.cuda(), causal=causal)
    path = nn.last_attention_path()
    ro, rl = O.naive_attention(q.double(), k.double(), v.double(), causal=causal, return_lse=True)
    nn.set_attention_path(1); o2, lse2 = nn._flash_attention(q.cuda(), k.cuda(), v.cuda(), causal=causal); nn.set_attention_path(0)
    print(f"B{B} H{QH}/{KH} {QL}x{KL} causal={causal}: path {path} | tensor-core err o {(o.double().cpu()-ro).abs().max():.2e} lse {(lse.double().cpu()-rl).abs().max():.2e} | SIMT err o {(o2.double().cpu()-ro).abs().max():.2e}", flush=True)
# backward
for (B, QH, KH, QL, KL, causal) in [(1, 1, 1, 128, 128, False), (1, 2, 2, 256, 256, False), (2, 2, 1, 255, 511, False), (1, 2, 2, 1024, 1024, True), (3, 4, 2, 513, 513, True)]:
    q = torch.randn(B, QH, QL, 64); k = torch.randn(B, KH, KL, 64); v = torch.randn(B, KH, KL, 64); dO = torch.randn(B, QH, QL, 64)
    qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
    o, lse = nn._flash_attention(qd, kd, vd, causal=causal)
    dq, dk, dv, _ = nn.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
    path = nn.last_attention_path()
    rq, rk, rv, _ = O.naive_attention_bwd(dO.double(), q.double(), k.double(), v.double(), causal=causal)
    e = lambda a, b: (a.double().cpu() - b).abs().max().item()
    print(f"bwd B{B} H{QH}/{KH} {QL}x{KL} causal={causal}: path {path} | err dq {e(dq, rq):.2e} dk {e(dk, rk):.2e} dv {e(dv, rv):.2e} (mag {rq.abs().max():.1f} {rk.abs().max():.1f} {rv.abs().max():.1f})", flush=True)
# C1 timing: f32 E=64 L=4096 H=4 B=4 non-causal
q, k, v = (torch.randn(4, 4, 4096, 64, device="cuda") for _ in range(3))
def T(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
f = 4.0 * 4 * 4 * 4096 * 4096 * 64
t1 = T(lambda: nn._flash_attention(q, k, v, causal=False))
nn.set_attention_path(1); t0 = T(lambda: nn._flash_attention(q, k, v, causal=False)); nn.set_attention_path(0)
print(f"C1 forward: tensor-core split path {t1:.3f} ms {f/t1/1e9:.1f} TF/s | SIMT {t0:.3f} ms {f/t0/1e9:.1f} TF/s")
dO = torch.randn_like(q)
o, lse = nn._flash_attention(q, k, v, causal=False)
tb1 = T(lambda: nn.grad_flash_attention(dO, o, lse, q, k, v, causal=False))
nn.set_attention_path(1); tb0 = T(lambda: nn.grad_flash_attention(dO, o, lse, q, k, v, causal=False)); nn.set_attention_path(0)
print(f"C1 backward: tensor-core split path {tb1:.3f} ms {2.5*f/tb1/1e9:.1f} TF/s | SIMT {tb0:.3f} ms {2.5*f/tb0/1e9:.1f} TF/s | fwd+bwd {3.5*f/(t1+tb1)/1e9:.1f} vs {3.5*f/(t0+tb0)/1e9:.1f} TF/s")
