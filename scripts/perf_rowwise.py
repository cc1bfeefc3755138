"""CUDA-event timing of the HBM-bound ops vs the measured copy bandwidth (development aid)."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch
import nnop_b200 as nn

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(f"{name:46s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {100*gbs/PEAK:5.1f}% of measured copy BW", flush=True)

def main():
    for dtype, s in ((torch.bfloat16, 2), (torch.float32, 4)):
        tag = str(dtype)[6:]
        for n, emb in ((65536, 4096), (8192, 4096), (1024, 1024)):
            x = torch.randn(n, emb, device="cuda", dtype=dtype); w = torch.rand(emb, device="cuda", dtype=dtype)
            b = torch.rand(emb, device="cuda", dtype=dtype); dy = torch.randn_like(x)
            y, rstd = nn._rms_norm(x, w)
            report(f"rms_norm fwd {tag} ({emb},{n})", 2*emb*n*s + emb*s + 4*n, timeit(lambda: nn._rms_norm(x, w)))
            report(f"rms_norm bwd {tag} ({emb},{n})", 3*emb*n*s + emb*s + 4*n + 4*emb, timeit(lambda: nn.grad_rms_norm(dy, rstd, x, w)))
            y, mu, rs = nn._layer_norm(x, w, b)
            report(f"layer_norm fwd {tag} ({emb},{n})", 2*emb*n*s + 2*emb*s + 8*n, timeit(lambda: nn._layer_norm(x, w, b)))
            report(f"layer_norm bwd {tag} ({emb},{n})", 3*emb*n*s + emb*s + 8*n + 2*emb*s, timeit(lambda: nn.grad_layer_norm(dy, mu, rs, x, w, b)))
        for cols, N in ((1024, 8192), (16384, 8192)):
            x = torch.randn(cols, N, device="cuda", dtype=dtype); dy = torch.randn_like(x)
            y = nn.online_softmax(x)
            report(f"softmax fwd {tag} ({N},{cols})", 2*N*cols*s, timeit(lambda: nn._softmax_fwd(x) if hasattr(nn, '_softmax_fwd') else nn.online_softmax(x)))
            report(f"softmax bwd {tag} ({N},{cols})", 3*N*cols*s, timeit(lambda: nn.grad_online_softmax(dy, y)))
        B, QH, KH, L, E = 8, 32, 8, 8192, 128
        q = torch.randn(B, QH, L, E, device="cuda", dtype=dtype); k = torch.randn(B, KH, L, E, device="cuda", dtype=dtype)
        pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
        cos, sin = nn.LlamaRotaryEmbedding(E)(pos); cos, sin = cos.cuda(), sin.cuda()
        report(f"llama_rope {tag} q({E},{L},{QH},{B}) k(..{KH}..)", 2*(q.numel()+k.numel())*s + 2*(E//2)*L*B*4,
               timeit(lambda: nn.llama_rope(q, k, cos=cos, sin=sin)))
        del q, k

if __name__ == "__main__":
    main()
