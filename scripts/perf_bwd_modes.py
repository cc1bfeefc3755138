"""A/B timing of the backward kernel variants (nnop_set_bwd_pair_mode) inside one process:
2 = one CTA per tile, 3 = persistent, 4 = persistent CTA pairs exchanging dQ halves (DUO).  Interleaved repeats, CUDA events, backward only."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch
import nnop_b200 as nn


def run(B, H, KH, L, E, causal, modes=(2, 3), iters=5, rounds=3, dtype=torch.bfloat16):
    q = torch.randn(B, H, L, E, device="cuda", dtype=dtype)
    k = torch.randn(B, KH, L, E, device="cuda", dtype=dtype)
    v = torch.randn(B, KH, L, E, device="cuda", dtype=dtype)
    dO = torch.randn(B, H, L, E, device="cuda", dtype=dtype)
    f = 2.5 * 4.0 * B * H * L * L * E * (0.5 if causal else 1.0)
    o, lse = nn._flash_attention(q, k, v, causal=causal)
    best = {m: 1e9 for m in modes}
    for r in range(rounds):
        for m in modes:
            nn.set_bwd_pair_mode(m)
            nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal)
            e1.record()
            torch.cuda.synchronize()
            best[m] = min(best[m], e0.elapsed_time(e1) / iters)
    nn.set_bwd_pair_mode(0)
    print(f"B{B} H{H}/{KH} L{L} E{E} causal={causal}: " +
          "  ".join(f"mode {m}: {t:.3f} ms {f / t / 1e9:.0f} TF/s" for m, t in best.items()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:   # python scripts/perf_bwd_modes.py 3,4  -> compare these modes
        import functools
        run = functools.partial(run, modes=tuple(int(x) for x in sys.argv[1].split(",")))
    run(8, 32, 32, 8192, 128, True)
    run(4, 32, 8, 8192, 128, True)
    run(8, 32, 32, 2048, 128, True)
    run(2, 32, 32, 8192, 128, False)
    run(8, 32, 32, 8192, 64, True)
