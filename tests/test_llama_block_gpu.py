"""BASELINE config 3: the Llama-3-8B attention block as the reference's ops compose it --
rms_norm over the hidden dim, (stand-in) q/k/v projections, llama_rope, GQA causal flash attention
with 32 query / 8 kv heads of E=128 -- forward and backward through the autograd wrappers (the
ChainRules rrules' twin) vs the oracle composed the same way in fp64.  Projections are plain
matmuls (library GEMMs are not part of the hot path); hidden is reduced to keep the oracle fast."""
import pytest
import torch

from helpers import max_abs
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _block(ops, x, w_norm, wq, wk, wv, cos, sin, QH, KH, E, *, oracle):
    B, L, hidden = x.shape
    xn = ops["rms_norm"](x.reshape(B * L, hidden), w_norm).reshape(B, L, hidden)
    heads = lambda t, H: t.reshape(B, L, H, E).permute(0, 2, 1, 3).contiguous()
    q, k, v = heads(xn @ wq, QH), heads(xn @ wk, KH), heads(xn @ wv, KH)
    q, k = ops["rope"](q, k, cos, sin)
    return ops["attn"](q, k, v)


@pytest.mark.parametrize("L", [384, 1000])
def test_llama3_attention_block_fwd_bwd(nnop, L):
    B, hidden, QH, KH, E = 2, 512, 32, 8, 128
    g = torch.Generator().manual_seed(L)
    x = torch.randn(B, L, hidden, generator=g)
    w_norm = 1.0 + 0.1 * torch.randn(hidden, generator=g)
    wq = torch.randn(hidden, QH * E, generator=g) / hidden ** 0.5
    wk = torch.randn(hidden, KH * E, generator=g) / hidden ** 0.5
    wv = torch.randn(hidden, KH * E, generator=g) / hidden ** 0.5
    dO = torch.randn(B, QH, L, E, generator=g)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = nnop.LlamaRotaryEmbedding(E)(pos)

    bf = lambda t: t.to(torch.bfloat16).cuda()
    leaves = [bf(t).requires_grad_(True) for t in (x, w_norm, wq, wk, wv)]
    ours = dict(rms_norm=lambda a, w: nnop.rms_norm(a, w),
                rope=lambda q, k, c, s: nnop.llama_rope(q, k, cos=c, sin=s),
                attn=lambda q, k, v: nnop.flash_attention(q, k, v, causal=True))
    o = _block(ours, *leaves, cos.cuda(), sin.cuda(), QH, KH, E, oracle=False)
    assert nnop.last_attention_path() == 1
    grads = torch.autograd.grad(o, leaves, bf(dO))

    # oracle, fp64, from the SAME bf16-rounded leaves
    rl = [t.detach().double().cpu().requires_grad_(True) for t in leaves]
    ref = dict(rms_norm=lambda a, w: O.naive_rms_norm(a, w),
               rope=lambda q, k, c, s: O.naive_llama_rope(q, k, cos=c, sin=s),
               attn=lambda q, k, v: O.naive_attention(q, k, v, causal=True))
    ro = _block(ref, *rl, cos.double(), sin.double(), QH, KH, E, oracle=True)
    rgrads = torch.autograd.grad(ro, rl, bf(dO).double().cpu())

    assert max_abs(o, ro) < 3e-2, "o"          # chain of four bf16 ops: 1.5x the single-op bound
    for name, a, b in zip(("dx", "dw_norm", "dwq", "dwk", "dwv"), grads, rgrads):
        rel = (a.double().cpu() - b).norm().item() / max(b.norm().item(), 1e-12)
        assert rel < 2e-2, (name, rel)         # reference-style norm-wise check (isapprox, rtol)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-2), (torch.float32, 2e-5)])
def test_config_c3_rowwise_ops_full_shape(nnop, dtype, tol):
    """BASELINE config C3 at its full size, one batch element: rms_norm / layer_norm over hidden 4096 on
    L = 8192 rows (forward and backward), llama_rope on q (32 heads) / k (8 heads) of E = 128 at L = 8192 (forward
    and pullback), online_softmax at the reference's benchmark shape (8192 x 1024, benchmarks/main.jl:279-300):
    every element against the fp64 oracle.  These are the shapes `bench.py`'s `secondary` block times."""
    hidden, L, QH, KH, E = 4096, 8192, 32, 8, 128
    g = torch.Generator().manual_seed(33)
    x = torch.randn(L, hidden, generator=g).to(dtype)
    dy = torch.randn(L, hidden, generator=g).to(dtype)
    w = (1.0 + 0.1 * torch.randn(hidden, generator=g)).to(dtype)
    b = torch.rand(hidden, generator=g).to(dtype)
    xd, dyd, wd, bd = x.cuda(), dy.cuda(), w.cuda(), b.cuda()
    X, DY, W, Bb = x.double(), dy.double(), w.double(), b.double()
    y, rstd = nnop._rms_norm(xd, wd)
    assert max_abs(y, O.naive_rms_norm(X, W)) < tol
    dx, dw = nnop.grad_rms_norm(dyd, rstd, xd, wd)
    rx, rw = O.naive_rms_norm_bwd(DY, X, W)
    assert max_abs(dx, rx) < tol
    assert max_abs(dw, rw) < 1e-3 * max(1.0, rw.abs().max().item())      # dw stays Float32: 8192 fp32 adds per column
    y, mean, rstd = nnop._layer_norm(xd, wd, bd)
    assert max_abs(y, O.naive_layer_norm(X, W, Bb)) < tol
    dx, dw, db = nnop.grad_layer_norm(dyd, mean, rstd, xd, wd, bd)
    rx, rw, rb = O.naive_layer_norm_bwd(DY, X, W, Bb)
    assert max_abs(dx, rx) < tol
    wtol = 2e-2 if dtype == torch.bfloat16 else 1e-3                      # dw / db are returned in T
    assert max_abs(dw, rw) < wtol * max(1.0, rw.abs().max().item())
    assert max_abs(db, rb) < wtol * max(1.0, rb.abs().max().item())
    del xd, dyd, X, DY, rx, y, dx
    # RoPE
    q = torch.randn(1, QH, L, E, generator=g).to(dtype)
    k = torch.randn(1, KH, L, E, generator=g).to(dtype)
    pos = torch.arange(L, dtype=torch.float32).view(1, L)
    cos, sin = nnop.LlamaRotaryEmbedding(E)(pos)
    q1, k1 = nnop.llama_rope(q.cuda(), k.cuda(), cos=cos.cuda(), sin=sin.cuda())
    q2, k2 = O.naive_llama_rope(q.double(), k.double(), cos=cos.double(), sin=sin.double())
    assert max_abs(q1, q2) < 4 * tol and max_abs(k1, k2) < 4 * tol
    qb, kb = nnop.grad_llama_rope(q.cuda(), k.cuda(), cos=cos.cuda(), sin=sin.cuda())
    q3, k3 = O.naive_llama_rope(q.double(), k.double(), cos=cos.double(), sin=sin.double(), bwd=True)
    assert max_abs(qb, q3) < 4 * tol and max_abs(kb, k3) < 4 * tol
    # softmax, reference benchmark shape
    xs = torch.randn(1024, 8192, generator=g).to(dtype)
    ds = torch.randn(1024, 8192, generator=g).to(dtype)
    s = nnop.online_softmax(xs.cuda())
    assert max_abs(s, O.naive_softmax(xs.double())) < (2e-3 if dtype == torch.bfloat16 else 1e-6)
    gs = nnop.grad_online_softmax(ds.cuda(), s)
    assert max_abs(gs, O.naive_softmax_bwd(ds.double(), s.double().cpu())) < tol
