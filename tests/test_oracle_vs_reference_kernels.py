"""Pins the oracle from a second side (CPU, no GPU): `oracle/ref_kernels.py` restates the reference's FUSED
KERNELS tile loop by tile loop (src/attention.jl:1-131, src/attention_bwd.jl:1-197, src/softmax.jl:1-58,
src/rms_norm.jl:3-115, src/layer_norm.jl:8-148, src/rope/llama_rope.jl:24-65); `oracle/oracle.py` restates the
NAIVE functions its tests compare them with.  The reference's own test suite asserts exactly this equality on a
GPU; here both sides run in float64 on the reference's shapes (ragged tiles on both axes, GQA, causal, key padding
mask, pair), so that a misreading of either file -- head mapping, mask alignment, bounds guards, the Δ / l
preprocess, the layout of `pair` -- shows up as a mismatch.  Also pinned: the new library's single residual
lse = ms + log(ls) and its fully normalised O are what the reference's (o, ms, ls) carry."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_kernels as RK


def _jl(t):
    """row-major torch (B, H, L, E) -> the reference's column-major (E, L, H, B) as a NumPy array."""
    return t.permute(*reversed(range(t.dim()))).contiguous().numpy()


def _case(B, QH, KH, QL, KL, E, seed, pair, mask):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, QH, QL, E, generator=g, dtype=torch.float64)
    k = torch.randn(B, KH, KL, E, generator=g, dtype=torch.float64)
    v = torch.randn(B, KH, KL, E, generator=g, dtype=torch.float64)
    dO = torch.randn(B, QH, QL, E, generator=g, dtype=torch.float64)
    pr = torch.randn(B, KL, QL, QH, generator=g, dtype=torch.float64) if pair else None
    m = None
    if mask:   # test/attention_tests.jl:27-28: the tail of the last batch element is padding
        m = torch.ones(B, KL, dtype=torch.bool)
        m[-1, -5:] = False
    return q, k, v, dO, pr, m


@pytest.mark.parametrize("gsz", [16, 32])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("mask", [False, True])
def test_fused_attention_kernels_equal_naive_attention(gsz, causal, pair, mask):
    # (key lengths leave >= 8 keys in the last tile: with the 5 padded keys of `mask` no tile is fully masked --
    # the reference's own shapes have the same property; see the NaN test below)
    for (B, QH, KH, QL, KL) in [(2, 2, 2, 40, 40), (2, 4, 2, 64, 64), (1, 6, 2, 47, 47), (2, 2, 1, 31, 56), (1, 2, 2, 64, 24)]:
        if causal and QL != KL:
            continue
        E = 16
        q, k, v, dO, pr, m = _case(B, QH, KH, QL, KL, E, 7 * QL + KL + gsz, pair, mask)
        prj = None if pr is None else _jl(pr)            # (B, KL, QL, QH) -> (QH, QL, KL, B)
        mj = None if m is None else m.numpy().T          # (B, KL) -> (KL, B)
        o, ms, ls = RK.flash_attention_fwd(_jl(q), _jl(k), _jl(v), prj, mj, causal=causal, gsz=gsz)
        ro, rl = O.naive_attention(q, k, v, pr, causal=causal, kpad_mask=m, return_lse=True)
        assert np.abs(o - _jl(ro)).max() < 1e-12
        # the new library's residual: lse = m + log(l)
        assert np.abs(ms + np.log(ls) - _jl(rl)).max() < 1e-12
        dq, dk, dv, dp = RK.flash_attention_bwd(_jl(dO), o, ms, ls, _jl(q), _jl(k), _jl(v), prj, mj, causal=causal,
                                                gsz=gsz)
        rq, rk, rv, rp = O.naive_attention_bwd(dO, q, k, v, pr, causal=causal, kpad_mask=m)
        assert np.abs(dq - _jl(rq)).max() < 1e-11
        assert np.abs(dk - _jl(rk)).max() < 1e-11
        assert np.abs(dv - _jl(rv)).max() < 1e-11
        if pair:
            assert np.abs(dp - _jl(rp)).max() < 1e-11


def test_reference_kernel_gives_nan_on_a_fully_masked_tile():
    """src/attention.jl:91: a key tile in which every in-range key is masked has m_ij = -Inf and
    exp(-Inf - (-Inf)) = NaN, which poisons the row for good.  The new library defines such rows / tiles instead
    (masked keys contribute exactly 0; a row without any key gives 0 and lse = -inf): SURVEY.md App. C 4."""
    q, k, v, _, _, m = _case(1, 2, 2, 33, 33, 16, 5, False, True)      # KL = 33, gsz = 16: the last tile holds only key 32
    o, ms, ls = RK.flash_attention_fwd(_jl(q), _jl(k), _jl(v), None, m.numpy().T, causal=False, gsz=16)
    assert np.isnan(o).all()
    ro = O.naive_attention(q, k, v, causal=False, kpad_mask=m)            # the naive function (and the new kernels) are finite
    assert torch.isfinite(ro).all()


def test_reference_backward_float16_staging_is_a_1e3_effect():
    """The reference rounds Q and K to Float16 in its backward for every T (src/attention_bwd.jl:19-20).  The new
    library does not; this measures what that rounding costs the reference itself against
    the exact gradients: of the order of its own test tolerance (atol = rtol = 1e-3, test/attention_tests.jl:42-48)."""
    q, k, v, dO, _, _ = _case(2, 2, 2, 64, 64, 16, 3, False, False)
    o, ms, ls = RK.flash_attention_fwd(_jl(q), _jl(k), _jl(v), causal=False, gsz=32)
    exact = RK.flash_attention_bwd(_jl(dO), o, ms, ls, _jl(q), _jl(k), _jl(v), causal=False, gsz=32)
    staged = RK.flash_attention_bwd(_jl(dO), o, ms, ls, _jl(q), _jl(k), _jl(v), causal=False, gsz=32, stage_f16=True)
    errs = [np.abs(a - b).max() for a, b in zip(exact[:3], staged[:3])]
    assert 1e-5 < max(errs) < 2e-2


@pytest.mark.parametrize("N", [32, 33, 255, 513])      # test/softmax_tests.jl:12
def test_online_softmax_kernel_equals_naive(N):
    x = torch.rand(4, N, generator=torch.Generator().manual_seed(N), dtype=torch.float64)   # torch (cols, N)
    y = RK.online_softmax(x.numpy().T, gsz=64)          # reference (N, cols)
    assert np.abs(y.T - O.naive_softmax(x).numpy()).max() < 1e-14


@pytest.mark.parametrize("emb,n", [(15, 1), (255, 4), (257, 17), (512, 23)])   # test/rmsnorm_tests.jl:11-14
@pytest.mark.parametrize("offset", [0.0, 1.0])
def test_norm_kernels_equal_naive(emb, n, offset):
    g = torch.Generator().manual_seed(emb + n)
    x = torch.rand(n, emb, generator=g, dtype=torch.float64).requires_grad_(True)    # torch (n, emb)
    w = torch.rand(emb, generator=g, dtype=torch.float64).requires_grad_(True)
    b = torch.rand(emb, generator=g, dtype=torch.float64).requires_grad_(True)
    dy = torch.rand(n, emb, generator=g, dtype=torch.float64)
    xj, dyj = x.detach().numpy().T, dy.numpy().T
    # RMS norm
    y, rstd = RK.rms_norm_fwd(xj, w.detach().numpy(), eps=1e-6, offset=offset)
    ry = O.naive_rms_norm(x, w, eps=1e-6, offset=offset)
    assert np.abs(y.T - ry.detach().numpy()).max() < 1e-13
    rdx, rdw = torch.autograd.grad(ry, (x, w), dy)
    dx, dw = RK.rms_norm_bwd(dyj, rstd, xj, w.detach().numpy(), offset=offset)
    assert np.abs(dx.T - rdx.numpy()).max() < 1e-12 and np.abs(dw - rdw.numpy()).max() < 1e-11
    # layer norm
    y, mu, rs = RK.layer_norm_fwd(xj, w.detach().numpy(), b.detach().numpy(), eps=1e-6)
    ry = O.naive_layer_norm(x, w, b, eps=1e-6)
    assert np.abs(y.T - ry.detach().numpy()).max() < 1e-12
    rdx, rdw, rdb = torch.autograd.grad(ry, (x, w, b), dy)
    dx, dw, db = RK.layer_norm_bwd(dyj, mu, rs, xj, w.detach().numpy())
    assert np.abs(dx.T - rdx.numpy()).max() < 1e-11
    assert np.abs(dw - rdw.numpy()).max() < 1e-10 and np.abs(db - rdb.numpy()).max() < 1e-11


@pytest.mark.parametrize("L,QH,KH", [(13, 1, 3), (257, 4, 1), (64, 5, 5)])     # test/rope_tests.jl:21-24
def test_rope_kernel_equals_naive(L, QH, KH):
    E, B = 16, 2
    g = torch.Generator().manual_seed(L)
    q = torch.randn(B, QH, L, E, generator=g, dtype=torch.float64)
    k = torch.randn(B, KH, L, E, generator=g, dtype=torch.float64)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = O.llama_rotary_embedding(E, pos)
    rq, rk = O.naive_llama_rope(q, k, cos=cos.double(), sin=sin.double())
    cj, sj = _jl(cos.double()), _jl(sin.double())        # torch (B, L, E) -> (E, L, B)
    qo, ko = RK.llama_rope(_jl(q), _jl(k), cj, sj)
    assert np.abs(qo - _jl(rq)).max() < 1e-13 and np.abs(ko - _jl(rk)).max() < 1e-13
    # backward = the same kernel with sin -> -sin (src/rope/llama_rope.jl:86): it inverts the rotation
    qb, kb = RK.llama_rope(qo, ko, cj, sj, sin_sign=-1.0)
    # (up to cos^2 + sin^2 = 1 in the Float32 tables the reference builds on the host, :15-22)
    assert np.abs(qb - _jl(q)).max() < 1e-6 and np.abs(kb - _jl(k)).max() < 1e-6
