"""A/B timing of the CTA-pair backward vs the single-CTA backward inside one process (development aid)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
def T(fn, n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (B, H, KH, L, causal) in [(8, 32, 32, 8192, True), (4, 32, 32, 8192, False), (4, 32, 8, 8192, True), (8, 32, 32, 2048, True)]:
    q = torch.randn(B, H, L, 128, device="cuda", dtype=torch.bfloat16); dO = torch.randn_like(q)
    k = torch.randn(B, KH, L, 128, device="cuda", dtype=torch.bfloat16); v = torch.randn_like(k)
    o, lse = nn._flash_attention(q, k, v, causal=causal)
    f = 2.5 * 4.0 * B * H * L * L * 128 * (0.5 if causal else 1.0)
    res = []
    for rep in range(2):
        for mode in (0, 1):
            nn.set_bwd_pair_mode(mode)
            t = T(lambda: nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal))
            res.append(f"{'pair' if mode else 'single'} {t:.3f} ms {f/t/1e9:.0f} TF/s")
    print(f"B{B} H{H}/{KH} L{L} causal={causal}: " + " | ".join(res), flush=True)
