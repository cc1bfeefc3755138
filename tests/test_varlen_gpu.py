"""Packed variable-length flash attention (additive API, SURVEY.md 8 f1 / BASELINE config 4) through
the C ABI vs the oracle applied sequence by sequence.  Tolerance: BASELINE.json's 2e-2 max-abs for
BF16/FP16 on O, dQ, dK, dV (scaled by magnitude above 2, as in test_attention_gpu.py)."""
import math

import pytest
import torch

from helpers import kernel_err, max_abs
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _packed(lens_q, lens_k, QH, KH, E, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    TQ, TK = sum(lens_q), sum(lens_k)
    q = torch.randn(QH, TQ, E, generator=g).to(dtype)
    k = torch.randn(KH, TK, E, generator=g).to(dtype)
    v = torch.randn(KH, TK, E, generator=g).to(dtype)
    dO = torch.randn(QH, TQ, E, generator=g).to(dtype)
    cu = lambda ls: torch.tensor([0] + list(torch.tensor(ls).cumsum(0)), dtype=torch.int32)
    return q, k, v, dO, cu(lens_q), cu(lens_k)


def _oracle(q, k, v, dO, cu_q, cu_k, causal):
    """Per-sequence naive attention fwd + bwd (fp64), re-packed."""
    o = torch.zeros_like(q, dtype=torch.float64)
    lse = torch.zeros(q.shape[0], q.shape[1], dtype=torch.float64)
    dq = torch.zeros_like(q, dtype=torch.float64)
    dk = torch.zeros_like(k, dtype=torch.float64)
    dv = torch.zeros_like(v, dtype=torch.float64)
    for z in range(cu_q.numel() - 1):
        a, b = int(cu_q[z]), int(cu_q[z + 1])
        c, d = int(cu_k[z]), int(cu_k[z + 1])
        if b == a:
            continue
        if d == c:
            lse[:, a:b] = -math.inf
            continue
        qs, ks, vs, ds = (t.double()[None] for t in (q[:, a:b], k[:, c:d], v[:, c:d], dO[:, a:b]))
        oz, lz = O.naive_attention(qs, ks, vs, causal=causal, return_lse=True)
        gq, gk, gv, _ = O.naive_attention_bwd(ds, qs, ks, vs, causal=causal)
        o[:, a:b], lse[:, a:b] = oz[0], lz[0]
        dq[:, a:b], dk[:, c:d], dv[:, c:d] = gq[0], gk[0], gv[0]
    return o, lse, dq, dk, dv


def _check(nnop, lens_q, lens_k, QH, KH, E, dtype, causal, seed=0):
    q, k, v, dO, cu_q, cu_k = _packed(lens_q, lens_k, QH, KH, E, dtype, seed)
    dev = lambda t: t.cuda()
    mq, mk = max(lens_q), max(lens_k)
    o, lse = nnop._flash_attention_varlen(dev(q), dev(k), dev(v), dev(cu_q), dev(cu_k), mq, mk, causal=causal)
    assert nnop.last_attention_path() == 1
    dq, dk, dv = nnop.grad_flash_attention_varlen(dev(dO), o, lse, dev(q), dev(k), dev(v), dev(cu_q),
                                                  dev(cu_k), mq, mk, causal=causal)
    ro, rl, rq, rk, rv = _oracle(q, k, v, dO, cu_q, cu_k, causal)
    assert max_abs(o, ro) < TOL, "o"
    fin = torch.isfinite(rl)
    assert torch.equal(torch.isfinite(lse.cpu()), fin), "lse finiteness"
    assert max_abs(lse.cpu()[fin], rl[fin]) < 1e-3, "lse"
    assert kernel_err(dq, rq) < TOL, "dq"   # error beyond the output's rounding to T (helpers.kernel_err)
    assert kernel_err(dk, rk) < TOL, "dk"
    assert kernel_err(dv, rv) < TOL, "dv"


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("E", [64, 128])
def test_varlen_ragged(nnop, E, dtype, causal):
    # lengths straddling the 128 / 256 tile edges, a 1-token sequence and the reference's ragged Ls
    lens = [255, 1, 128, 513, 256, 129, 64, 511]
    _check(nnop, lens, lens, 4, 4, E, dtype, causal)


@pytest.mark.parametrize("causal", [False, True])
def test_varlen_gqa(nnop, causal):
    lens = [300, 77, 1024, 5]
    _check(nnop, lens, lens, 8, 2, 128, torch.bfloat16, causal)


def test_varlen_cross_lengths(nnop):
    # non-causal with different query / key lengths per sequence (test/attention_tests.jl:14-18 QL != KL)
    _check(nnop, [100, 257, 31], [513, 64, 200], 2, 2, 128, torch.bfloat16, False)


def test_varlen_empty_sequences(nnop):
    # zero-length sequences in the middle and at the end; a sequence with queries but no keys
    _check(nnop, [130, 0, 64, 0], [130, 0, 64, 0], 2, 1, 64, torch.bfloat16, True)
    _check(nnop, [40, 200], [0, 200], 2, 2, 128, torch.bfloat16, False)


def test_varlen_matches_dense(nnop):
    """A packed batch of equal-length sequences must reproduce the dense kernel bit for bit."""
    B, H, L, E = 3, 4, 384, 128
    g = torch.Generator().manual_seed(3)
    q, k, v, dO = (torch.randn(B, H, L, E, generator=g).to(torch.bfloat16).cuda() for _ in range(4))
    o, lse = nnop._flash_attention(q, k, v, causal=True)
    dq, dk, dv, _ = nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
    pk = lambda t: t.permute(1, 0, 2, 3).reshape(H, B * L, E).contiguous()
    cu = torch.arange(0, (B + 1) * L, L, dtype=torch.int32).cuda()
    o2, lse2 = nnop._flash_attention_varlen(pk(q), pk(k), pk(v), cu, cu, L, L, causal=True)
    d2 = nnop.grad_flash_attention_varlen(pk(dO), o2, lse2, pk(q), pk(k), pk(v), cu, cu, L, L, causal=True)
    assert torch.equal(o2, pk(o))
    assert torch.equal(lse2, lse.permute(1, 0, 2).reshape(H, B * L))
    # dQ goes through fp32 atomics whose order is not fixed: equal up to rounding of the last bit
    for a, b in zip(d2, (dq, dk, dv)):
        assert max_abs(a, pk(b)) <= 2 ** -7 * max(1.0, b.abs().max().item())


def test_varlen_config4_shape_properties(nnop):
    """BASELINE config 4 shape: 64 sequences, L log-uniform in [128, 16384], H=32, E=128, bf16
    causal.  Size-independent checks: causal row 0 of every sequence = v[0]; constant V => constant O;
    sum_k dV = sum_q dO per sequence and head."""
    g = torch.Generator().manual_seed(0)
    lens = torch.exp(torch.empty(64).uniform_(math.log(128), math.log(16384), generator=g)).round().int().tolist()
    H, E = 32, 128
    T = sum(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    q = torch.randn(H, T, E, generator=g).to(torch.bfloat16).cuda()
    k = torch.randn(H, T, E, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(H, T, E, generator=g).to(torch.bfloat16).cuda()
    o, lse = nnop._flash_attention_varlen(q, k, v, cu, cu, max(lens), max(lens), causal=True)
    starts = cu[:-1].long()
    assert torch.equal(o[:, starts], v[:, starts])          # first row sees only its own key
    qk0 = (q[:, starts].float() * k[:, starts].float()).sum(-1) / math.sqrt(E)
    assert (lse[:, starts] - qk0).abs().max().item() < 1e-3     # one visible key: lse = q.k/sqrt(E)
    vc = torch.full_like(v, 0.5)
    oc, _ = nnop._flash_attention_varlen(q, k, vc, cu, cu, max(lens), max(lens), causal=True)
    assert (oc.float() - 0.5).abs().max().item() < 4e-3
    dO = torch.randn(H, T, E, generator=g).to(torch.bfloat16).cuda()
    dq, dk, dv = nnop.grad_flash_attention_varlen(dO, o, lse, q, k, v, cu, cu, max(lens), max(lens), causal=True)
    for z in (0, 17, 63):
        a, b = int(cu[z]), int(cu[z + 1])
        s_dv = dv[:, a:b].float().sum(1)
        s_do = dO[:, a:b].float().sum(1)
        assert (s_dv - s_do).abs().max().item() < 2e-2 * max(1.0, s_do.abs().max().item())
    # every sequence, O and all three gradients, against an independent GPU implementation (torch SDPA + autograd on
    # the sequence's own slice): catches a sequence / tile of the full packed grid going wrong; bound 2e-2 + 2 ulps
    import torch.nn.functional as F
    from helpers import ulp_T
    for z in range(64):
        a, b = int(cu[z]), int(cu[z + 1])
        qz, kz, vz = (t[None, :, a:b].clone().requires_grad_(True) for t in (q, k, v))
        oz = F.scaled_dot_product_attention(qz, kz, vz, is_causal=True)
        gq, gk, gv = torch.autograd.grad(oz, (qz, kz, vz), dO[None, :, a:b])
        for got, ref, what in ((o, oz.detach(), "o"), (dq, gq, "dq"), (dk, gk, "dk"), (dv, gv, "dv")):
            d = (got[:, a:b].float() - ref[0].float()).abs()
            assert not (d > TOL + 2 * ulp_T(ref[0], torch.bfloat16).float()).any(), (what, z, lens[z], d.max().item())
    # one short sequence against the oracle
    z = min(range(64), key=lambda i: lens[i])
    a, b = int(cu[z]), int(cu[z + 1])
    qs, ks, vs, ds = (t[:2, a:b].double().cpu()[None] for t in (q, k, v, dO))
    ro = O.naive_attention(qs, ks, vs, causal=True)
    rq, rk, rv, _ = O.naive_attention_bwd(ds, qs, ks, vs, causal=True)
    assert max_abs(o[:2, a:b], ro[0]) < TOL
    assert kernel_err(dq[:2, a:b], rq[0]) < TOL
    assert kernel_err(dk[:2, a:b], rk[0]) < TOL
    assert kernel_err(dv[:2, a:b], rv[0]) < TOL


def test_varlen_errors(nnop):
    q = torch.randn(2, 64, 128, device="cuda")
    cu = torch.tensor([0, 64], dtype=torch.int32, device="cuda")
    with pytest.raises(nnop.NNopError, match="Float16 / BFloat16"):
        nnop._flash_attention_varlen(q, q, q, cu, cu, 64, 64, causal=False)
    qh = torch.randn(2, 64, 32, device="cuda").bfloat16()
    with pytest.raises(nnop.NNopError, match="64 and 128"):
        nnop._flash_attention_varlen(qh, qh, qh, cu, cu, 64, 64, causal=False)


@pytest.mark.parametrize("mode", [3, 101, 103])
def test_varlen_persistent_backward(nnop, mode):
    """The persistent backward on packed batches (dynamic tile queue over (sequence, kv head, kv block),
    sequence found by binary search in a device-built tile prefix): oracle parity on every scenario
    of this file, and dK / dV bit-identical to the one-CTA-per-tile kernel.  Modes 101 / 103 squeeze
    the queue onto 1 / 3 CTAs so that tiles of different sequences follow each other in one CTA."""
    cases = [([255, 1, 128, 513, 256, 129, 64, 511], None, 4, 4, 128, True),
             ([255, 1, 128, 513, 256, 129, 64, 511], None, 4, 4, 64, False),
             ([300, 77, 1024, 5], None, 8, 2, 128, True),
             ([100, 257, 31], [513, 64, 200], 2, 2, 128, False),
             ([130, 0, 64, 0], None, 2, 1, 64, True),
             ([40, 200], [0, 200], 2, 2, 128, False),
             ([0, 200, 64], [300, 200, 64], 2, 2, 128, False)]   # keys without queries: dK = dV = 0
    try:
        for lens_q, lens_k, QH, KH, E, causal in cases:
            lens_k = lens_k or lens_q
            nnop.set_bwd_pair_mode(mode)
            _check(nnop, lens_q, lens_k, QH, KH, E, torch.bfloat16, causal, seed=len(lens_q))
            q, k, v, dO, cu_q, cu_k = (t.cuda() for t in _packed(lens_q, lens_k, QH, KH, E, torch.bfloat16, 11))
            mq, mk = max(lens_q), max(lens_k)
            o, lse = nnop._flash_attention_varlen(q, k, v, cu_q, cu_k, mq, mk, causal=causal)
            got = nnop.grad_flash_attention_varlen(dO, o, lse, q, k, v, cu_q, cu_k, mq, mk, causal=causal)
            nnop.set_bwd_pair_mode(2)
            ref = nnop.grad_flash_attention_varlen(dO, o, lse, q, k, v, cu_q, cu_k, mq, mk, causal=causal)
            assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]), (lens_q, lens_k)
            assert max_abs(got[0], ref[0]) <= 2 ** -7 * max(1.0, ref[0].abs().max().item())
    finally:
        nnop.set_bwd_pair_mode(0)


def test_varlen_many_short_sequences(nnop):
    """A thousand short sequences (more sequences than threads in a CTA): exercises the chunked
    block-wide count of the forward's tile lookup and the chunked tile prefix + binary search of the
    persistent backward, both of which see one sequence per thread in the other tests."""
    import random
    rng = random.Random(5)
    lens = [rng.choice([0, 1, 7, 40, 128, 129, 200, 300]) for _ in range(1000)]
    try:
        for mode in (0, 3, 107):
            nnop.set_bwd_pair_mode(mode)
            _check(nnop, lens, lens, 4, 2, 64, torch.bfloat16, True, seed=9)
    finally:
        nnop.set_bwd_pair_mode(0)


@pytest.mark.parametrize("mode", [2, 101, 103])
def test_varlen_persistent_forward(nnop, mode):
    """The persistent forward on packed batches (sequence found in a per-CTA coarse tile prefix, partial
    last tiles stored row by row, key-less sequences zero-filled by the scheduler warp): O and lse bit
    for bit against the one-CTA-per-tile kernel, which the other tests of this file pin to the oracle."""
    import random
    rng = random.Random(3)
    cases = [([255, 1, 128, 513, 256, 129, 64, 511], None, 4, 4, 128, True),
             ([255, 1, 128, 513, 256, 129, 64, 511], None, 4, 4, 64, False),
             ([300, 77, 1024, 5], None, 8, 2, 128, True),
             ([100, 257, 31], [513, 64, 200], 2, 2, 128, False),
             ([130, 0, 64, 0], None, 2, 1, 64, True),
             ([40, 200], [0, 200], 2, 2, 128, False),
             ([rng.choice([0, 1, 7, 40, 128, 129, 200, 300, 600]) for _ in range(300)], None, 2, 2, 64, True)]
    try:
        for lens_q, lens_k, QH, KH, E, causal in cases:
            lens_k = lens_k or lens_q
            q, k, v, dO, cu_q, cu_k = (t.cuda() for t in _packed(lens_q, lens_k, QH, KH, E, torch.bfloat16, 13))
            mq, mk = max(lens_q), max(lens_k)
            nnop.set_fwd_mode(1)
            o_ref, lse_ref = nnop._flash_attention_varlen(q, k, v, cu_q, cu_k, mq, mk, causal=causal)
            nnop.set_fwd_mode(mode)
            for _ in range(2):
                o, lse = nnop._flash_attention_varlen(q, k, v, cu_q, cu_k, mq, mk, causal=causal)
                assert torch.equal(o, o_ref), (lens_q[:8], lens_k[:8], E, causal)
                assert torch.equal(lse, lse_ref), (lens_q[:8], lens_k[:8], E, causal)
    finally:
        nnop.set_fwd_mode(0)
