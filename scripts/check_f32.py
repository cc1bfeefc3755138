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
