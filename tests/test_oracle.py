"""Pins the CPU oracle: against independent torch implementations, torch autograd on the
fp64 graph, analytic known-answer tests built from the reference tests' own inputs, and the
committed golden vectors.  Runs without a GPU."""
import math

import pytest
import torch

from helpers import load_golden, max_abs
from oracle import oracle as O

f64 = torch.float64


def _rand_attn(B, QH, KH, QL, KL, E, seed=0, pair=False, mask=False):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, QH, QL, E, generator=g, dtype=f64)
    k = torch.randn(B, KH, KL, E, generator=g, dtype=f64)
    v = torch.randn(B, KH, KL, E, generator=g, dtype=f64)
    dO = torch.randn(B, QH, QL, E, generator=g, dtype=f64)
    pr = torch.randn(B, KL, QL, QH, generator=g, dtype=f64) if pair else None
    m = None
    if mask:
        m = torch.ones(B, KL, dtype=torch.bool)
        m[-1, -11:] = False
    return q, k, v, dO, pr, m


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("g", [1, 2])
def test_attention_vs_sdpa(causal, g):
    """independent implementation: torch SDPA (math backend on CPU), incl. GQA expansion."""
    q, k, v, _, _, _ = _rand_attn(2, 4, 4 // g, 40, 40, 32)
    o = O.naive_attention(q, k, v, causal=causal)
    kk = k.repeat_interleave(g, dim=1)
    vv = v.repeat_interleave(g, dim=1)
    ref = torch.nn.functional.scaled_dot_product_attention(q, kk, vv, is_causal=causal)
    assert max_abs(o, ref) < 1e-12


def test_attention_pair_and_mask_vs_sdpa():
    q, k, v, _, pair, mask = _rand_attn(2, 2, 2, 31, 47, 16, pair=True, mask=True)
    o = O.naive_attention(q, k, v, pair, causal=False, kpad_mask=mask)
    bias = pair.permute(0, 3, 2, 1).clone()
    bias = bias.masked_fill(~mask.view(2, 1, 1, 47), float("-inf"))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias)
    assert max_abs(o, ref) < 1e-12


@pytest.mark.parametrize("causal,pair,mask,g", [(False, False, False, 1), (True, False, False, 2),
                                                (False, True, True, 1), (True, True, True, 3)])
def test_attention_bwd_vs_autograd(causal, pair, mask, g):
    QH = 6 if g == 3 else 4
    q, k, v, dO, pr, m = _rand_attn(2, QH, QH // g, 33, 33, 16, seed=3, pair=pair, mask=mask)
    leaves = [t.clone().requires_grad_(True) for t in (q, k, v)]
    prl = pr.clone().requires_grad_(True) if pr is not None else None
    o = O.naive_attention(*leaves, prl, causal=causal, kpad_mask=m)
    grads = torch.autograd.grad(o, leaves + ([prl] if prl is not None else []), dO)
    dq, dk, dv, dpair = O.naive_attention_bwd(dO, q, k, v, pr, causal=causal, kpad_mask=m)
    assert max_abs(dq, grads[0]) < 1e-11
    assert max_abs(dk, grads[1]) < 1e-11
    assert max_abs(dv, grads[2]) < 1e-11
    if pr is not None:
        assert max_abs(dpair, grads[3]) < 1e-11


def test_attention_lse_and_masked_rows():
    q, k, v, _, _, _ = _rand_attn(1, 1, 1, 8, 8, 16)
    o, lse = O.naive_attention(q, k, v, causal=True, return_lse=True)
    s = torch.einsum("bhqe,bhke->bhqk", q, k) / math.sqrt(16)
    s = s.masked_fill(torch.ones(8, 8).triu(1).bool(), float("-inf"))
    assert max_abs(lse, torch.logsumexp(s, dim=-1)) < 1e-12
    # fully masked row: reference behaviour is NaN, the kernels' documented behaviour is 0
    mask = torch.zeros(1, 8, dtype=torch.bool)
    assert torch.isnan(O.naive_attention(q, k, v, causal=False, kpad_mask=mask)).all()
    o0, l0 = O.naive_attention(q, k, v, causal=False, kpad_mask=mask, zero_masked_rows=True,
                               return_lse=True)
    assert (o0 == 0).all() and torch.isinf(l0).all()


def test_softmax_rms_ln_vs_torch():
    g = torch.Generator().manual_seed(5)
    x = torch.rand(7, 257, generator=g, dtype=f64)
    w = torch.rand(257, generator=g, dtype=f64)
    b = torch.rand(257, generator=g, dtype=f64)
    dy = torch.randn(7, 257, generator=g, dtype=f64)
    assert max_abs(O.naive_softmax(x), torch.softmax(x, -1)) < 1e-14
    assert max_abs(O.naive_layer_norm(x, w, b), torch.nn.functional.layer_norm(x, (257,), w, b, 1e-6)) < 1e-12
    assert max_abs(O.naive_rms_norm(x, w), torch.nn.functional.rms_norm(x, (257,), w, 1e-6)) < 1e-12
    # closed-form backward vs autograd
    for off in (0.0, 1.0):
        xl, wl = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        gx, gw = torch.autograd.grad(O.naive_rms_norm(xl, wl, offset=off), (xl, wl), dy)
        dx, dw = O.naive_rms_norm_bwd(dy, x, w, offset=off)
        assert max_abs(dx, gx) < 1e-11 and max_abs(dw, gw) < 1e-11
    xl, wl, bl = (t.clone().requires_grad_(True) for t in (x, w, b))
    gx, gw, gb = torch.autograd.grad(O.naive_layer_norm(xl, wl, bl), (xl, wl, bl), dy)
    dx, dw, db = O.naive_layer_norm_bwd(dy, x, w, b)
    assert max_abs(dx, gx) < 1e-10 and max_abs(dw, gw) < 1e-11 and max_abs(db, gb) < 1e-11
    xl = x.clone().requires_grad_(True)
    y = O.naive_softmax(xl)
    (gx,) = torch.autograd.grad(y, xl, dy)
    assert max_abs(O.naive_softmax_bwd(dy, y.detach()), gx) < 1e-13


@pytest.mark.parametrize("L", [13, 255, 256, 257, 1024, 1025])
def test_rope_known_answer_from_reference_test(L):
    """test/rope_tests.jl:21-56 feeds q = k = ones, positions 0..L-1, dim 16.  Then analytically
    out[i] = cos(p f_i) - sin(p f_i) for i < 8 and cos + sin for i >= 8, f_i = 10000^(-2i/16);
    the gradient of sum(q') + sum(k') w.r.t. q is cos + sin (i < 8) and cos - sin (i >= 8)."""
    dim = 16
    pos = torch.arange(L, dtype=torch.float32).view(1, L)
    cos, sin = O.llama_rotary_embedding(dim, pos)
    assert cos.shape == (1, L, dim) and cos.dtype == torch.float32
    f = torch.tensor([10000.0 ** (-2.0 * i / dim) for i in range(dim // 2)], dtype=f64)
    ang = torch.arange(L, dtype=f64).view(L, 1) * f
    # table construction is Float32 in the reference: agree with the fp64 angle to fp32 accuracy
    assert max_abs(cos[0, :, :8], torch.cos(ang)) < 2e-4 * max(1.0, L / 1024)
    assert torch.equal(cos[..., :8], cos[..., 8:]) and torch.equal(sin[..., :8], sin[..., 8:])
    q = torch.ones(1, 3, L, dim, dtype=f64)
    k = torch.ones(1, 5, L, dim, dtype=f64)
    c, s = cos.double(), sin.double()
    qo, ko = O.naive_llama_rope(q, k, cos=c, sin=s)
    exp_lo = (c - s)[0, :, :8]
    exp_hi = (c + s)[0, :, 8:]
    for h in range(3):
        assert max_abs(qo[0, h, :, :8], exp_lo) < 1e-15 and max_abs(qo[0, h, :, 8:], exp_hi) < 1e-15
    assert max_abs(ko[0, 4, :, :8], exp_lo) < 1e-15
    gq, gk = O.naive_llama_rope(torch.ones_like(q), torch.ones_like(k), cos=c, sin=s, bwd=True)
    ql = q.clone().requires_grad_(True)
    kl = k.clone().requires_grad_(True)
    a, b_ = O.naive_llama_rope(ql, kl, cos=c, sin=s)
    ga, gb = torch.autograd.grad(a.sum() + b_.sum(), (ql, kl))
    assert max_abs(gq, ga) < 1e-15 and max_abs(gk, gb) < 1e-15
    assert max_abs(gq[0, 0, :, :8], (c + s)[0, :, :8]) < 1e-15


def test_uniform_softmax_known_answer():
    x = torch.full((4, 33), 0.25, dtype=f64)
    assert max_abs(O.naive_softmax(x), torch.full_like(x, 1 / 33)) < 1e-16


def test_oracle_matches_golden():
    att = load_golden("attention.npz")
    for name, d in att.items():
        causal = bool(d["causal"])
        o, lse = O.naive_attention(d["q"], d["k"], d["v"], d.get("pair"), causal=causal,
                                   kpad_mask=d.get("kpad_mask"), return_lse=True)
        dq, dk, dv, dpair = O.naive_attention_bwd(d["dO"], d["q"], d["k"], d["v"], d.get("pair"),
                                                  causal=causal, kpad_mask=d.get("kpad_mask"))
        for got, key in ((o, "o"), (lse, "lse"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
            assert max_abs(got, d[key]) < 1e-12, (name, key)
        if "pair" in d:
            assert max_abs(dpair, d["dpair"]) < 1e-12
    row = load_golden("rowwise.npz")
    for name, d in row.items():
        if name.startswith("softmax"):
            assert max_abs(O.naive_softmax(d["x"]), d["y"]) < 1e-14
        elif name.startswith("rms"):
            assert max_abs(O.naive_rms_norm(d["x"], d["w"], offset=float(d["offset"])), d["y"]) < 1e-13
        else:
            assert max_abs(O.naive_layer_norm(d["x"], d["w"], d["b"]), d["y"]) < 1e-12
    for name, d in load_golden("rope.npz").items():
        L, E = d["cos"].shape[1], d["cos"].shape[2]
        B = d["cos"].shape[0]
        pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
        cos, sin = O.llama_rotary_embedding(E, pos)
        assert torch.equal(cos, d["cos"]) and torch.equal(sin, d["sin"])
        qo, ko = O.naive_llama_rope(d["q"], d["k"], cos=cos.double(), sin=sin.double())
        assert max_abs(qo, d["q_out"]) < 1e-15 and max_abs(ko, d["k_out"]) < 1e-15
