"""llama_rope through the C ABI vs the oracle on the reference grid (test/rope_tests.jl:21-56)."""
import pytest
import torch

from helpers import load_golden, max_abs
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [13, 255, 256, 257, 1024, 1025])
def test_rope_reference_grid(nnop, L):
    dim, batch = 16, 1
    emb = nnop.LlamaRotaryEmbedding(dim)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(batch, 1)
    cos, sin = emb(pos)
    rc, rs = O.llama_rotary_embedding(dim, pos)
    assert torch.equal(cos, rc) and torch.equal(sin, rs)
    for QH in (1, 3, 4, 5):
        for KH in (1, 3, 4, 5):
            q = torch.ones(batch, QH, L, dim)
            k = torch.ones(batch, KH, L, dim)
            qd, kd = q.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
            q1, k1 = nnop.llama_rope(qd, kd, cos=cos.cuda(), sin=sin.cuda())
            q2, k2 = O.naive_llama_rope(q.double(), k.double(), cos=cos.double(), sin=sin.double())
            assert max_abs(q1, q2) < 1e-6 and max_abs(k1, k2) < 1e-6
            gq, gk = torch.autograd.grad(q1.sum() + k1.sum(), (qd, kd))
            rq, rk = O.naive_llama_rope(torch.ones_like(q2), torch.ones_like(k2), cos=cos.double(),
                                        sin=sin.double(), bwd=True)
            assert max_abs(gq, rq) < 1e-6 and max_abs(gk, rk) < 1e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-2), (torch.float16, 4e-3)])
def test_rope_random_and_roundtrip(nnop, dtype, tol):
    """random inputs, E=128 (Llama-3), GQA head counts; bwd(fwd(x)) == x (rotation is orthogonal)."""
    g = torch.Generator().manual_seed(0)
    B, QH, KH, L, E = 2, 8, 2, 301, 128
    q = torch.randn(B, QH, L, E, generator=g).to(dtype)
    k = torch.randn(B, KH, L, E, generator=g).to(dtype)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = nnop.LlamaRotaryEmbedding(E)(pos)
    q1, k1 = nnop.llama_rope(q.cuda(), k.cuda(), cos=cos.cuda(), sin=sin.cuda())
    q2, k2 = O.naive_llama_rope(q.double(), k.double(), cos=cos.double(), sin=sin.double())
    assert max_abs(q1, q2) < tol * 4 and max_abs(k1, k2) < tol * 4
    qb, kb = nnop.grad_llama_rope(q1, k1, cos=cos.cuda(), sin=sin.cuda())
    assert max_abs(qb, q) < tol * 8 and max_abs(kb, k) < tol * 8


def test_rope_golden_and_odd_head_dim(nnop):
    for name, d in load_golden("rope.npz").items():
        q1, k1 = nnop.llama_rope(d["q"].float().cuda(), d["k"].float().cuda(), cos=d["cos"].cuda(),
                                 sin=d["sin"].cuda())
        assert max_abs(q1, d["q_out"]) < 2e-6 and max_abs(k1, d["k_out"]) < 2e-6
        qb, kb = nnop.grad_llama_rope(d["q"].float().cuda(), d["k"].float().cuda(), cos=d["cos"].cuda(),
                                      sin=d["sin"].cuda())
        assert max_abs(qb, d["q_bwd"]) < 2e-6 and max_abs(kb, d["k_bwd"]) < 2e-6
    # E = 6: half dim 3 is not a vector multiple -> scalar path
    pos = torch.arange(7, dtype=torch.float32).view(1, 7)
    cos, sin = nnop.LlamaRotaryEmbedding(6)(pos)
    q = torch.randn(1, 2, 7, 6)
    k = torch.randn(1, 1, 7, 6)
    q1, k1 = nnop.llama_rope(q.cuda(), k.cuda(), cos=cos.cuda(), sin=sin.cuda())
    q2, k2 = O.naive_llama_rope(q.double(), k.double(), cos=cos.double(), sin=sin.double())
    assert max_abs(q1, q2) < 2e-6 and max_abs(k1, k2) < 2e-6
