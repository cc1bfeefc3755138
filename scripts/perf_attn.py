"""Quick CUDA-event timing of flash attention fwd / bwd (development aid; bench.py is the contract)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch
import nnop_b200 as nn

def run(B, H, KH, L, E, causal, dtype=torch.bfloat16, bwd=True, iters=5):
    q = torch.randn(B, H, L, E, device="cuda", dtype=dtype)
    k = torch.randn(B, KH, L, E, device="cuda", dtype=dtype)
    v = torch.randn(B, KH, L, E, device="cuda", dtype=dtype)
    dO = torch.randn(B, H, L, E, device="cuda", dtype=dtype)
    f = 4.0 * B * H * L * L * E * (0.5 if causal else 1.0)
    for _ in range(2):
        o, lse = nn._flash_attention(q, k, v, causal=causal)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        o, lse = nn._flash_attention(q, k, v, causal=causal)
    ev[1].record()
    torch.cuda.synchronize()
    tf = ev[0].elapsed_time(ev[1]) / iters
    msg = f"B{B} H{H}/{KH} L{L} E{E} causal={causal} {str(dtype)[6:]}: fwd {tf:.3f} ms {f/tf/1e9:.1f} TF/s path={nn.last_attention_path()}"
    if bwd:
        for _ in range(1):
            nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal)
        torch.cuda.synchronize()
        ev[2].record()
        for _ in range(iters):
            nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal)
        ev[3].record()
        torch.cuda.synchronize()
        tb = ev[2].elapsed_time(ev[3]) / iters
        msg += f" | bwd {tb:.3f} ms {2.5*f/tb/1e9:.1f} TF/s path={nn.last_attention_path()} | fwd+bwd {3.5*f/(tf+tb)/1e9:.1f} TF/s"
    print(msg, flush=True)

if __name__ == "__main__":
    bwd = "--nobwd" not in sys.argv
    run(8, 32, 32, 8192, 128, True, bwd=bwd)
    run(8, 32, 32, 8192, 128, False, bwd=False)
    run(4, 32, 8, 8192, 128, True, bwd=bwd)
    run(8, 32, 32, 2048, 128, True, bwd=bwd)
    run(8, 32, 32, 8192, 64, True, bwd=False)
    run(4, 4, 4, 4096, 64, False, dtype=torch.float32, bwd=bwd)
