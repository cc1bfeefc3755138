"""Float32 E = 128 forward on the tensor cores (one-tile split kernel) against the SIMT kernel; timing (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "nnop.jl_b200"))
import torch
import nnop_b200 as nn
torch.manual_seed(0)
bad = 0
for (B, QH, KH, QL, KL, causal, pad, sc) in [(2, 4, 4, 512, 512, True, False, 1.0), (1, 4, 2, 300, 517, False, False, 1.0),
                                             (3, 2, 2, 255, 1024, False, True, 1.0), (1, 2, 1, 129, 64, False, False, 1.0),
                                             (2, 2, 2, 1000, 1000, True, True, 1.0), (1, 1, 1, 1, 1, False, False, 1.0),
                                             (1, 2, 2, 1024, 1024, True, False, 1e3), (1, 2, 2, 1024, 1024, False, False, 1e-3)]:
    q = torch.randn(B, QH, QL, 128, device="cuda") * (sc if sc < 1 else 1); k = torch.randn(B, KH, KL, 128, device="cuda") * (sc if sc < 1 else 1)
    v = torch.randn(B, KH, KL, 128, device="cuda") * sc
    m = None
    if pad:
        m = torch.ones(B, KL, dtype=torch.bool, device="cuda"); m[-1, -11:] = False; m[0, 5:40] = False
    nn.set_attention_path(1); o0, l0 = nn._flash_attention(q, k, v, causal=causal, kpad_mask=m); p0 = nn.last_attention_path()
    nn.set_attention_path(0); o1, l1 = nn._flash_attention(q, k, v, causal=causal, kpad_mask=m); p1 = nn.last_attention_path()
    torch.cuda.synchronize()
    ref = max(1.0, o0.abs().max().item()) if sc != 1.0 else 1.0
    do = (o0 - o1).abs().max().item() / ref; dl = (l0 - l1).abs().max().item()
    ok = do <= 1e-4 and dl <= 1e-4 and p1 == 1 and p0 == 0
    bad += not ok
    print(f"B{B} H{QH}/{KH} QL{QL} KL{KL} causal={causal} pad={pad} x{sc}: paths {p0}->{p1} max|do| {do:.3e} max|dlse| {dl:.3e} {'ok' if ok else 'MISMATCH'}", flush=True)
print("mismatches:", bad)
def T(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (B, H, L, causal) in [(4, 4, 4096, False), (4, 16, 4096, True)]:
    q, k, v = (torch.randn(B, H, L, 128, device="cuda") for _ in range(3))
    f = 4.0 * B * H * L * L * 128 * (0.5 if causal else 1.0)
    for path in (1, 0):
        nn.set_attention_path(path)
        t = T(lambda: nn._flash_attention(q, k, v, causal=causal))
        print(f"f32 E=128 B{B} H{H} L{L} causal={causal} path {'SIMT' if path == 1 else 'tcgen05'}: {t:.3f} ms {f/t/1e9:.0f} TF/s", flush=True)
nn.set_attention_path(0)
