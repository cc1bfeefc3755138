#!/usr/bin/env python
"""The reference's own benchmark script (benchmarks/main.jl) replayed on B200 through the C ABI.

Same shapes, dtypes and flag combinations as the reference (Float32 everywhere):
  layer_norm / rms_norm  1024 x 1024                       (benchmarks/main.jl:70-187)
  llama_rope             dim 64, 3 heads, L 1024, B 4      (:189-261)
  online_softmax         8192 x 1024                       (:279-300)
  flash_attention        E 64, L 2048, H 4, B 4; causal x padmask x pair   (:305-386)
Like the reference's `@btime ... KA.synchronize` every sample is one call followed by a device
synchronisation ("call" column, host launch path of the Python twin included); the "gpu" column is the
device time of the same call replayed from a CUDA graph (no host in the loop).  The "naive" rows are the
un-fused definitions the reference benchmarks against (its test helpers), written with torch ops on
the same GPU.  Development aid: bench.py is the contract.
"""
import math
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import nnop_b200 as nn  # noqa: E402

dev = torch.device("cuda", 0)


def call_us(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e6


def gpu_us(fn, inner=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for _ in range(inner):
                fn()
    except Exception as e:  # an op that cannot be captured
        return float("nan")
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (5 * inner)


def row(name, fn):
    print(f"{name:58s} call {call_us(fn):9.1f} us   gpu {gpu_us(fn):9.1f} us", flush=True)


def grad_of(f, *xs):
    def run():
        ys = [x.detach().requires_grad_(True) for x in xs]
        f(*ys).sum().backward()
    return run


def main():
    torch.manual_seed(0)
    f32 = dict(device=dev, dtype=torch.float32)
    # ---- layer norm / rms norm, 1024 x 1024 --------------------------------------------
    x = torch.rand(1024, 1024, **f32); w = torch.rand(1024, **f32); b = torch.rand(1024, **f32)
    naive_ln = lambda x, w, b: (x - x.mean(-1, keepdim=True)) * torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-6) * w + b
    naive_rms = lambda x, w: x * torch.rsqrt((x * x).mean(-1, keepdim=True) + 1e-6) * w
    with torch.no_grad():
        row("layer_norm 1024x1024 FWD naive", lambda: naive_ln(x, w, b))
        row("layer_norm 1024x1024 FWD nnop", lambda: nn.layer_norm(x, w, b))
    row("layer_norm 1024x1024 FWD+BWD naive", grad_of(naive_ln, x, w, b))
    row("layer_norm 1024x1024 FWD+BWD nnop", grad_of(lambda x, w, b: nn.layer_norm(x, w, b), x, w, b))
    with torch.no_grad():
        row("rms_norm 1024x1024 FWD naive", lambda: naive_rms(x, w))
        row("rms_norm 1024x1024 FWD nnop", lambda: nn.rms_norm(x, w))
    row("rms_norm 1024x1024 FWD+BWD naive", grad_of(naive_rms, x, w))
    row("rms_norm 1024x1024 FWD+BWD nnop", grad_of(lambda x, w: nn.rms_norm(x, w), x, w))
    # ---- RoPE: dim 64, 3 heads, L 1024, B 4 ---------------------------------------------
    E, L, H, B = 64, 1024, 3, 4
    q = torch.randn(B, H, L, E, **f32); k = torch.randn(B, H, L, E, **f32)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = nn.LlamaRotaryEmbedding(E)(pos)
    cos, sin = cos.to(dev), sin.to(dev)

    def naive_rope(q, k):
        c, s = cos[:, None], sin[:, None]
        rot = lambda t: torch.cat((-t[..., E // 2:], t[..., :E // 2]), -1)
        return q * c + rot(q) * s, k * c + rot(k) * s
    with torch.no_grad():
        row("llama_rope E64 H3 L1024 B4 FWD naive", lambda: naive_rope(q, k))
        row("llama_rope E64 H3 L1024 B4 FWD nnop", lambda: nn.llama_rope(q, k, cos=cos, sin=sin))
    row("llama_rope E64 H3 L1024 B4 FWD+BWD naive", grad_of(lambda q, k: sum(t.sum() for t in naive_rope(q, k)), q, k))
    row("llama_rope E64 H3 L1024 B4 FWD+BWD nnop",
        grad_of(lambda q, k: sum(t.sum() for t in nn.llama_rope(q, k, cos=cos, sin=sin)), q, k))
    # ---- softmax 8192 x 1024 (softmax over the 8192 axis: rows of the (cols, N) view) ----
    xs = torch.rand(1024, 8192, **f32)
    with torch.no_grad():
        row("softmax 8192x1024 naive", lambda: torch.softmax(xs, -1))
        row("softmax 8192x1024 nnop online_softmax", lambda: nn.online_softmax(xs))
    # ---- flash attention: E 64, L 2048, H 4, B 4 ------------------------------------------
    E, L, H, B = 64, 2048, 4, 4
    fl = 4.0 * B * H * L * L * E
    for causal in (False, True):
        for use_padmask in (False, True):
            for use_pair in (False, True):
                q, k, v = (torch.randn(B, H, L, E, **f32) for _ in range(3))
                pm = None
                if use_padmask:
                    pm = torch.ones(B, L, dtype=torch.bool, device=dev)
                    pm[-1, -11:] = False
                pair = torch.randn(B, L, L, H, **f32) if use_pair else None   # (QH,QL,KL,B) column-major

                def naive(q, k, v, pair=None):
                    s = (q @ k.transpose(-1, -2)) / math.sqrt(E)
                    if causal:
                        s = s.masked_fill(~torch.ones(L, L, dtype=torch.bool, device=dev).tril(), float("-inf"))
                    if pm is not None:
                        s = s + torch.log(pm.float())[:, None, None, :]
                    if pair is not None:
                        s = s + pair.permute(0, 3, 2, 1)
                    return torch.softmax(s, -1) @ v
                tag = f"attention causal={int(causal)} padmask={int(use_padmask)} pair={int(use_pair)}"
                args = (q, k, v) + ((pair,) if use_pair else ())
                ours = lambda q, k, v, pair=None: nn.flash_attention(q, k, v, pair, causal=causal, kpad_mask=pm)
                with torch.no_grad():
                    row(tag + " FWD naive", lambda: naive(*args))
                    row(tag + " FWD nnop", lambda: ours(*args))
                path = nn.last_attention_path()
                row(tag + " FWD+BWD naive", grad_of(naive, *args))
                row(tag + f" FWD+BWD nnop (path {'tcgen05' if path else 'SIMT'})", grad_of(ours, *args))


if __name__ == "__main__":
    main()
