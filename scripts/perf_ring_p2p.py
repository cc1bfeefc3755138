"""Config C5 through the C-ABI ring (nnop_ring_attn_fwd / _bwd): one process, one rank per visible GPU,
K / V pulled and gradient partials pushed over NVLink P2P.   python scripts/perf_ring_p2p.py [L] [iters]
Timing: CUDA events on every device's stream around the call, max over devices (never wall clock)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch  # noqa: E402
import nnop_b200 as nn  # noqa: E402


def timed(fn, devs, iters):
    for d in devs:
        torch.cuda.synchronize(d)
    ev = []
    for d in devs:
        with torch.cuda.device(d):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(torch.cuda.current_stream(d))
            ev.append((a, b))
    t0 = time.perf_counter()
    out = None
    for _ in range(iters):
        out = fn()
    for d, (a, b) in zip(devs, ev):
        with torch.cuda.device(d):
            b.record(torch.cuda.current_stream(d))
    for d in devs:
        torch.cuda.synchronize(d)
    wall = (time.perf_counter() - t0) / iters * 1e3
    return max(a.elapsed_time(b) for a, b in ev) / iters, wall, out


def main():
    W = torch.cuda.device_count()
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    B, H, E = 1, 32, 128
    Ll = L // W
    devs = list(range(W))
    mk = lambda d: torch.randn(B, H, Ll, E, device=f"cuda:{d}", dtype=torch.float32).to(torch.bfloat16)
    qs, ks, vs, dOs = ([mk(d) for d in devs] for _ in range(4))
    for causal in (True,):
        fwd = lambda: nn.p2p_ring_attention_forward(qs, ks, vs, causal=causal)
        os_, lses = fwd()
        bwd = lambda: nn.p2p_ring_attention_backward(dOs, os_, lses, qs, ks, vs, causal=causal)
        bwd()
        t_f, w_f, _ = timed(fwd, devs, iters)
        t_b, w_b, _ = timed(bwd, devs, iters)
        f_fwd = 4.0 * B * H * L * L * E * (0.5 if causal else 1.0)
        f_bwd = 2.5 * f_fwd
        tot = t_f + t_b
        print(f"ring p2p (C ABI) causal={causal} W={W} L={L} H={H} E={E} bf16: fwd {t_f:.2f} ms (host wall {w_f:.2f}), "
              f"bwd {t_b:.2f} ms (host wall {w_b:.2f})")
        print(f"  fwd {f_fwd / t_f / 1e9:.0f} TFLOP/s, bwd {f_bwd / t_b / 1e9:.0f}, fwd+bwd {(f_fwd + f_bwd) / tot / 1e9:.0f} "
              f"TFLOP/s whole job = {(f_fwd + f_bwd) / tot / 1e9 / W:.0f} per GPU", flush=True)


if __name__ == "__main__":
    main()
