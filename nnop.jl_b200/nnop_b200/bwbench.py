"""Device-side timing of the HBM-bound ops (online_softmax, rms_norm, layer_norm, llama_rope; SURVEY.md
section 8(d) "secondary" metric: GB/s against the measured copy bandwidth).

Method: every op is called through the C ABI on pre-allocated buffers; `nsets` independent buffer sets are
rotated so that the bytes touched between two uses of one set exceed the 126 MB L2 (a re-run on warm
L2 lines would report L2, not HBM, bandwidth); the launches are captured into ONE CUDA graph and the graph
replay is timed with CUDA events, which removes the host's ctypes / Python launch cost (18-20 us per call,
more than the kernel itself for the reference's benchmark shapes) from the figure.  Algorithmic bytes are
SURVEY.md section 8(d)'s formulas.
"""
from __future__ import annotations

import math

import torch

from ._lib import lib, check
from .ops import LlamaRotaryEmbedding, _dt

L2_BYTES = 126 << 20


def _graph_ms(launch, nsets: int, min_launches: int = 24, replays: int = 5) -> float:
    """Mean device time (ms) of one `launch(i)` inside a replayed graph of >= min_launches launches."""
    k = max(min_launches, 2 * nsets)
    k = (k + nsets - 1) // nsets * nsets
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(nsets):
            launch(i)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(k):
            launch(i % nsets)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (replays * k)


def _nsets(bytes_per_set: int) -> int:
    return max(1, min(16, math.ceil(2.5 * L2_BYTES / max(bytes_per_set, 1))))


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return t.data_ptr()


def norm_ops(dtype, n: int, emb: int, which=("rms_norm", "layer_norm")):
    """-> list of dicts {op, shape, dtype, us, gbs, bytes} for fwd and bwd of the norms on x (emb, n)."""
    s = torch.empty((), dtype=dtype).element_size()
    dev = "cuda"
    ns = _nsets(3 * n * emb * s)
    xs = [torch.rand(n, emb, device=dev, dtype=torch.float32).to(dtype) for _ in range(ns)]
    dys = [torch.randn(n, emb, device=dev, dtype=torch.float32).to(dtype) for _ in range(ns)]
    ys = [torch.empty_like(x) for x in xs]
    w = torch.rand(emb, device=dev, dtype=torch.float32).to(dtype)
    b = torch.rand(emb, device=dev, dtype=torch.float32).to(dtype)
    mean = torch.empty(n, device=dev, dtype=torch.float32)
    rstd = torch.empty(n, device=dev, dtype=torch.float32)
    dwf = torch.empty(emb, device=dev, dtype=torch.float32)
    dw, db = torch.empty_like(w), torch.empty_like(w)
    nbytes = lib.nnop_norm_bwd_workspace_bytes(emb, n)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    dt = _dt(xs[0])
    out = []

    def rec(op, nb, ms):
        out.append(dict(op=op, shape=f"({emb},{n})", dtype=str(dtype)[6:], us=ms * 1e3, bytes=nb,
                        gbs=nb / ms / 1e6))

    if "rms_norm" in which:
        rec("rms_norm fwd", 2 * emb * n * s + emb * s + 4 * n,
            _graph_ms(lambda i: check(lib.nnop_rms_norm_fwd(_p(ys[i]), _p(rstd), _p(xs[i]), _p(w), dt, emb, n,
                                                            1e-6, 0.0, _st())), ns))
        rec("rms_norm bwd", 3 * emb * n * s + emb * s + 4 * n + 4 * emb,
            _graph_ms(lambda i: check(lib.nnop_rms_norm_bwd(_p(ys[i]), _p(dwf), _p(dys[i]), _p(rstd), _p(xs[i]),
                                                            _p(w), dt, emb, n, 0.0, _p(ws), nbytes, _st())), ns))
    if "layer_norm" in which:
        rec("layer_norm fwd", 2 * emb * n * s + 2 * emb * s + 8 * n,
            _graph_ms(lambda i: check(lib.nnop_layer_norm_fwd(_p(ys[i]), _p(mean), _p(rstd), _p(xs[i]), _p(w), _p(b),
                                                              dt, emb, n, 1e-6, _st())), ns))
        rec("layer_norm bwd", 3 * emb * n * s + emb * s + 8 * n + 2 * emb * s,
            _graph_ms(lambda i: check(lib.nnop_layer_norm_bwd(_p(ys[i]), _p(dw), _p(db), _p(dys[i]), _p(mean),
                                                              _p(rstd), _p(xs[i]), _p(w), dt, emb, n, _p(ws), nbytes,
                                                              _st())), ns))
    return out


def softmax_ops(dtype, cols: int, N: int):
    """online_softmax over dim 1 of x (N, cols): `cols` rows of N elements in memory."""
    s = torch.empty((), dtype=dtype).element_size()
    ns = _nsets(3 * N * cols * s)
    xs = [torch.randn(cols, N, device="cuda", dtype=torch.float32).to(dtype) for _ in range(ns)]
    dys = [torch.randn(cols, N, device="cuda", dtype=torch.float32).to(dtype) for _ in range(ns)]
    ys = [torch.empty_like(x) for x in xs]
    dxs = [torch.empty_like(x) for x in xs]
    dt = _dt(xs[0])
    out = []
    ms = _graph_ms(lambda i: check(lib.nnop_softmax_fwd(_p(ys[i]), _p(xs[i]), dt, N, cols, _st())), ns)
    out.append(dict(op="softmax fwd", shape=f"({N},{cols})", dtype=str(dtype)[6:], us=ms * 1e3,
                    bytes=2 * N * cols * s, gbs=2 * N * cols * s / ms / 1e6))
    ms = _graph_ms(lambda i: check(lib.nnop_softmax_bwd(_p(dxs[i]), _p(dys[i]), _p(ys[i]), dt, N, cols, _st())), ns)
    out.append(dict(op="softmax bwd", shape=f"({N},{cols})", dtype=str(dtype)[6:], us=ms * 1e3,
                    bytes=3 * N * cols * s, gbs=3 * N * cols * s / ms / 1e6))
    return out


def rope_op(dtype, B: int, QH: int, KH: int, L: int, E: int):
    s = torch.empty((), dtype=dtype).element_size()
    per = 2 * (B * QH * L * E + B * KH * L * E) * s
    ns = _nsets(per)
    qs = [torch.randn(B, QH, L, E, device="cuda", dtype=torch.float32).to(dtype) for _ in range(ns)]
    ks = [torch.randn(B, KH, L, E, device="cuda", dtype=torch.float32).to(dtype) for _ in range(ns)]
    qo = [torch.empty_like(t) for t in qs]
    ko = [torch.empty_like(t) for t in ks]
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = (t.cuda() for t in LlamaRotaryEmbedding(E)(pos))
    dt = _dt(qs[0])
    ms = _graph_ms(lambda i: check(lib.nnop_llama_rope(_p(qo[i]), _p(ko[i]), _p(qs[i]), _p(ks[i]), _p(cos), _p(sin),
                                                       dt, E, L, QH, KH, B, 1.0, _st())), ns)
    nb = per + 2 * (E // 2) * L * B * 4
    return [dict(op="llama_rope", shape=f"q({E},{L},{QH},{B}) k(..{KH}..)", dtype=str(dtype)[6:], us=ms * 1e3,
                 bytes=nb, gbs=nb / ms / 1e6)]


def secondary_block(peak_gbs: float, batch: int = 1):
    """The four bandwidth ops at BASELINE config C3's shapes (Llama-3-8B block, L = 8192, hidden 4096, bf16):
    rms_norm / layer_norm on x (4096, 8192*batch), llama_rope on q (128, 8192, 32, batch) + k (.., 8, ..),
    online_softmax on the reference's benchmark shape (8192, 1024) (benchmarks/main.jl:279-300)."""
    rows = []
    rows += norm_ops(torch.bfloat16, 8192 * batch, 4096)
    rows += rope_op(torch.bfloat16, batch, 32, 8, 8192, 128)
    rows += softmax_ops(torch.bfloat16, 1024, 8192)
    torch.cuda.empty_cache()
    return [dict(op=r["op"], shape=r["shape"], dtype=r["dtype"], us=round(r["us"], 2), gbs=round(r["gbs"], 1),
                 frac=round(r["gbs"] / peak_gbs, 4)) for r in rows]
