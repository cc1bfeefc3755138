"""Forward QUAD variant (nnop_set_fwd_mode(3): two softmax warps per 32 rows) against the default kernel:
results on a grid of shapes, then timing on C2 / C3 (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "nnop.jl_b200"))
import torch
import nnop_b200 as nn

torch.manual_seed(0)
def run(mode, *a, **kw):
    nn.set_fwd_mode(mode)
    try:
        return nn._flash_attention(*a, **kw)
    finally:
        nn.set_fwd_mode(0)

bad = 0
for (B, QH, KH, QL, KL, causal, dt, pad) in [
        (2, 4, 4, 2048, 2048, True, torch.bfloat16, False), (2, 4, 4, 2048, 2048, False, torch.bfloat16, False),
        (1, 4, 2, 300, 517, False, torch.float16, False), (1, 2, 2, 517, 517, True, torch.bfloat16, False),
        (2, 8, 2, 1024, 1024, True, torch.float16, False), (3, 2, 2, 255, 1024, False, torch.bfloat16, True),
        (1, 2, 1, 129, 64, False, torch.bfloat16, False), (2, 2, 2, 1000, 1000, True, torch.bfloat16, True),
        (1, 1, 1, 4096, 4096, True, torch.bfloat16, False)]:
    q = torch.randn(B, QH, QL, 128, device="cuda", dtype=dt) * 2
    k = torch.randn(B, KH, KL, 128, device="cuda", dtype=dt) * 2
    v = torch.randn(B, KH, KL, 128, device="cuda", dtype=dt)
    m = None
    if pad:
        m = torch.ones(B, KL, dtype=torch.bool, device="cuda"); m[-1, -11:] = False; m[0, 5:40] = False
    o1, l1 = run(1, q, k, v, causal=causal, kpad_mask=m)
    o3, l3 = run(3, q, k, v, causal=causal, kpad_mask=m)
    torch.cuda.synchronize()
    do = (o1.float() - o3.float()).abs().max().item()
    fin = torch.isfinite(l1)
    dl = (l1[fin] - l3[fin]).abs().max().item() if fin.any() else 0.0
    same_inf = bool((torch.isfinite(l3) == fin).all())
    ok = do <= 2e-2 and dl <= 1e-5 * max(1.0, l1[fin].abs().max().item()) and same_inf and not torch.isnan(o3).any()
    bad += not ok
    print(f"B{B} H{QH}/{KH} QL{QL} KL{KL} causal={causal} {dt} pad={pad}: max|do| {do:.3e} max|dlse| {dl:.3e} identical o: {bool((o1 == o3).all())} {'ok' if ok else 'MISMATCH'}", flush=True)
print("mismatches:", bad)

def T(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (B, QH, KH, L, causal) in [(8, 32, 32, 8192, True), (8, 32, 32, 8192, False), (4, 32, 8, 8192, True), (8, 32, 32, 2048, True)]:
    q = torch.randn(B, QH, L, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, KH, L, 128, device="cuda", dtype=torch.bfloat16); v = torch.randn_like(k)
    f = 4.0 * B * QH * L * L * 128 * (0.5 if causal else 1.0)
    for rep in range(2):
        for mode in (1, 3):
            nn.set_fwd_mode(mode)
            t = T(lambda: nn._flash_attention(q, k, v, causal=causal))
            print(f"B{B} H{QH}/{KH} L{L} causal={causal} mode {mode}: {t:.3f} ms {f/t/1e9:.0f} TF/s", flush=True)
nn.set_fwd_mode(0)
