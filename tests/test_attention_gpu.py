"""flash_attention forward + backward through the C ABI vs the oracle.

Grids are the reference's own (test/attention_tests.jl:6-49, test/causal_attention_tests.jl:6-47,
test/gqa_attention_tests.jl:6-34: Float32, E in {16,32,64}, ragged L, pair, kpad_mask, GQA), checked
element-wise at BASELINE.json's tolerance (max abs err <= 1e-4 for FP32, <= 2e-2 for BF16/FP16 on
O, dQ, dK, dV) -- stricter than the reference's norm-wise atol=rtol=1e-3 -- plus the 16-bit /
E=128 rows the reference leaves as TODO, golden vectors, and size-independent properties at the
full BASELINE config C2 shape."""
import math

import pytest
import torch

from helpers import kernel_err, load_golden, max_abs, reference_isapprox, ulp_T
from oracle import oracle as O

pytestmark = pytest.mark.gpu
F32_TOL = 1e-4
H16_TOL = 2e-2
LS = (255, 256, 511, 512, 1024)


def _inputs(B, QH, KH, QL, KL, E, dtype, seed, pair=False, mask=False):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, QH, QL, E, generator=g).to(dtype)
    k = torch.randn(B, KH, KL, E, generator=g).to(dtype)
    v = torch.randn(B, KH, KL, E, generator=g).to(dtype)
    dO = torch.randn(B, QH, QL, E, generator=g).to(dtype)
    pr = torch.randn(B, KL, QL, QH, generator=g).to(dtype) if pair else None
    m = None
    if mask:  # test/attention_tests.jl:27-28
        m = torch.ones(B, KL, dtype=torch.bool)
        m[-1, -11:] = False
    return q, k, v, dO, pr, m


def _check(nnop, q, k, v, dO, pr, m, causal, tol, expect_path=None):
    dev = lambda t: None if t is None else t.cuda()
    o, lse = nnop._flash_attention(dev(q), dev(k), dev(v), dev(pr), causal=causal, kpad_mask=dev(m))
    if expect_path is not None:
        assert nnop.last_attention_path() == expect_path
    D = lambda t: None if t is None else t.double()
    o_ref, lse_ref = O.naive_attention(D(q), D(k), D(v), D(pr), causal=causal, kpad_mask=m, return_lse=True)
    assert max_abs(o, o_ref) < tol, "o"
    assert max_abs(lse, lse_ref) < max(tol, 1e-4), "lse"
    # the reference's own assertion: isapprox(sum(o1), sum(o2); atol=1e-3, rtol=1e-3)
    assert abs(o.double().sum().item() - o_ref.sum().item()) <= max(1e-3, 1e-3 * abs(o_ref.sum().item())) \
        or tol > 1e-3
    dq, dk, dv, dpair = nnop.grad_flash_attention(dev(dO), o, lse, dev(q), dev(k), dev(v), dev(pr),
                                                  causal=causal, kpad_mask=dev(m))
    # the pullback is a function of (dO, o, lse, q, k, v): for 16-bit T the oracle takes delta = rowsum(dO o)
    # from the same (rounded) o the kernel was handed, as the reference's preprocess kernel does
    # (src/attention_bwd.jl:182-196); Float32 keeps the exact o
    o_arg = o.double().cpu() if tol > 1e-3 else None
    rq, rk, rv, rp = O.naive_attention_bwd(D(dO), D(q), D(k), D(v), D(pr), causal=causal, kpad_mask=m, o=o_arg)
    # BASELINE.json's bound (1e-4 Float32, 2e-2 16-bit) on the kernel's error: for 16-bit outputs that is the
    # error beyond the final rounding of the result to T (helpers.kernel_err).  Float32
    # keeps the plain max-abs 1e-4, also on the tensor-core path (E = 64): its dK / dV accumulators are
    # flushed with exact fp32 adds every few q blocks so that the bound holds for gradients of magnitude ~10.
    assert kernel_err(dq, rq) < tol, "dq"
    assert kernel_err(dk, rk) < tol, "dk"
    assert kernel_err(dv, rv) < tol, "dv"
    if pr is not None:
        assert kernel_err(dpair, rp) < tol, "dpair"
    return o, lse


# ----------------------------------------------------------------------------- reference grids
@pytest.mark.parametrize("E", [16, 32, 64])
@pytest.mark.parametrize("use_pair", [False, True])
@pytest.mark.parametrize("use_padmask", [False, True])
def test_noncausal_reference_grid(nnop, E, use_pair, use_padmask):
    for QL in LS:
        for KL in LS:
            q, k, v, dO, pr, m = _inputs(3, 2, 2, QL, KL, E, torch.float32, QL * 7 + KL, use_pair, use_padmask)
            # Float32 E <= 64 takes the split-operand tensor-core kernels (rows narrower than 64 are zero-padded
            # into the same operand tiles), with or without a pair bias
            _check(nnop, q, k, v, dO, pr, m, False, F32_TOL, expect_path=1)


@pytest.mark.parametrize("E", [16, 32, 64])
@pytest.mark.parametrize("use_pair", [False, True])
@pytest.mark.parametrize("use_padmask", [False, True])
def test_causal_reference_grid(nnop, E, use_pair, use_padmask):
    for L in LS:
        q, k, v, dO, pr, m = _inputs(3, 2, 2, L, L, E, torch.float32, L, use_pair, use_padmask)
        _check(nnop, q, k, v, dO, pr, m, True, F32_TOL, expect_path=1)


@pytest.mark.parametrize("QH", [4, 6, 8])
@pytest.mark.parametrize("KVH", [1, 2])
@pytest.mark.parametrize("causal", [False, True])
def test_gqa_reference_grid(nnop, QH, KVH, causal):
    for E in (32, 64):
        for L in (255, 256, 257, 512):
            q, k, v, dO, pr, m = _inputs(2, QH, KVH, L, L, E, torch.float32, L + QH)
            _check(nnop, q, k, v, dO, pr, m, causal, F32_TOL, expect_path=1)


def test_attention_golden(nnop):
    for name, d in load_golden("attention.npz").items():
        f = lambda key: d[key].float().cuda() if key in d else None
        mask = d["kpad_mask"].cuda() if "kpad_mask" in d else None
        causal = bool(d["causal"])
        o, lse = nnop._flash_attention(f("q"), f("k"), f("v"), f("pair"), causal=causal, kpad_mask=mask)
        assert max_abs(o, d["o"]) < F32_TOL and max_abs(lse, d["lse"]) < F32_TOL, name
        dq, dk, dv, dpair = nnop.grad_flash_attention(f("dO"), o, lse, f("q"), f("k"), f("v"), f("pair"),
                                                      causal=causal, kpad_mask=mask)
        assert max_abs(dq, d["dq"]) < F32_TOL and max_abs(dk, d["dk"]) < F32_TOL and max_abs(dv, d["dv"]) < F32_TOL, name
        if "pair" in d:
            assert max_abs(dpair, d["dpair"]) < F32_TOL


def test_autograd_wrapper_and_sum_gradient(nnop):
    """Zygote.gradient((q,k,v)->sum(flash_attention(...))) as in test/attention_tests.jl:35-41."""
    q, k, v, _, _, _ = _inputs(3, 2, 2, 255, 511, 32, torch.float32, 1)
    leaves = [t.cuda().requires_grad_(True) for t in (q, k, v)]
    o = nnop.flash_attention(*leaves, causal=False)
    grads = torch.autograd.grad(o.sum(), leaves)
    rq, rk, rv, _ = O.naive_attention_bwd(torch.ones(3, 2, 255, 32, dtype=torch.float64), q.double(), k.double(),
                                          v.double(), causal=False)
    for got, ref in zip(grads, (rq, rk, rv)):
        assert reference_isapprox(got, ref, 1e-3, 1e-3) and max_abs(got, ref) < F32_TOL


def test_fully_masked_rows_are_zero_not_nan(nnop):
    q, k, v, dO, _, _ = _inputs(2, 2, 2, 40, 40, 16, torch.float32, 3)
    m = torch.ones(2, 40, dtype=torch.bool)
    m[1, :] = False
    o, lse = nnop._flash_attention(q.cuda(), k.cuda(), v.cuda(), causal=False, kpad_mask=m.cuda())
    assert torch.isfinite(o).all() and (o[1] == 0).all() and torch.isinf(lse[1]).all()
    o_ref = O.naive_attention(q.double(), k.double(), v.double(), causal=False, kpad_mask=m, zero_masked_rows=True)
    assert max_abs(o, o_ref) < F32_TOL
    dq, dk, dv, _ = nnop.grad_flash_attention(dO.cuda(), o, lse, q.cuda(), k.cuda(), v.cuda(), causal=False,
                                              kpad_mask=m.cuda())
    rq, rk, rv, _ = O.naive_attention_bwd(dO.double(), q.double(), k.double(), v.double(), causal=False,
                                          kpad_mask=m, zero_masked_rows=True)
    assert max_abs(dq, rq) < F32_TOL and max_abs(dk, rk) < F32_TOL and max_abs(dv, rv) < F32_TOL


def test_error_behaviour(nnop):
    q = torch.randn(1, 2, 8, 16, device="cuda")
    with pytest.raises(nnop.NNopError, match="Embedding dim of Q"):
        nnop.flash_attention(q, torch.randn(1, 2, 8, 32, device="cuda"), torch.randn(1, 2, 8, 32, device="cuda"), causal=False)
    with pytest.raises(nnop.NNopError, match="Shapes of K"):
        nnop.flash_attention(q, torch.randn(1, 2, 8, 16, device="cuda"), torch.randn(1, 2, 9, 16, device="cuda"), causal=False)
    q48 = torch.randn(1, 2, 8, 48, device="cuda")
    with pytest.raises(nnop.NNopError, match="power-of-2"):
        nnop.flash_attention(q48, q48, q48, causal=False)
    q3 = torch.randn(1, 3, 8, 16, device="cuda")
    with pytest.raises(nnop.NNopError, match="divisible by number of KV heads"):
        nnop.flash_attention(q3, q[:, :2], q[:, :2], causal=False)
    with pytest.raises(nnop.NNopError, match="CUDA tensors only"):
        nnop.flash_attention(q.cpu(), q.cpu(), q.cpu(), causal=False)


# ----------------------------------------------------------------------------- tcgen05 path
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("E", [128, 64])
@pytest.mark.parametrize("causal", [False, True])
def test_tcgen05_path_vs_oracle(nnop, dtype, E, causal):
    for (B, QH, KH, QL, KL) in [(2, 2, 2, 128, 128), (1, 2, 2, 256, 256), (2, 4, 2, 255, 255),
                                (1, 2, 1, 257, 257), (1, 2, 2, 384, 384), (1, 3, 3, 511, 511),
                                (1, 2, 2, 1000, 1000), (1, 4, 1, 1024, 1024), (1, 2, 2, 200, 700),
                                (1, 2, 2, 700, 200), (1, 1, 1, 1, 1), (1, 1, 1, 130, 3)]:
        if causal and QL != KL:
            continue
        q, k, v, dO, _, _ = _inputs(B, QH, KH, QL, KL, E, dtype, QL + KL + E)
        try:
            _check(nnop, q, k, v, dO, None, None, causal, H16_TOL, expect_path=1)
        except AssertionError as e:
            raise AssertionError(f"shape {(B, QH, KH, QL, KL)}: {e}") from e


@pytest.mark.parametrize("causal", [False, True])
def test_tcgen05_matches_generic_path(nnop, causal):
    q, k, v, dO, _, _ = _inputs(2, 4, 2, 777, 777, 128, torch.bfloat16, 11)
    qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
    try:
        nnop.set_attention_path(1)
        o_g, lse_g = nnop._flash_attention(qd, kd, vd, causal=causal)
        g_g = nnop.grad_flash_attention(dOd, o_g, lse_g, qd, kd, vd, causal=causal)
        assert nnop.last_attention_path() == 0
        nnop.set_attention_path(2)
        o_f, lse_f = nnop._flash_attention(qd, kd, vd, causal=causal)
        assert nnop.last_attention_path() == 1
    finally:
        nnop.set_attention_path(0)
    assert max_abs(o_f, o_g) < H16_TOL and max_abs(lse_f, lse_g) < 1e-3
    g_f = nnop.grad_flash_attention(dOd, o_f, lse_f, qd, kd, vd, causal=causal)
    for a, b in zip(g_f[:3], g_g[:3]):  # two bf16 results: the bound applies beyond one output ulp
        assert kernel_err(a, b) < H16_TOL


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("E", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
def test_16bit_pair_on_tensor_cores(nnop, dtype, E, causal):
    """`pair` (src/attention.jl:55-62) with 16-bit inputs runs on the tcgen05 path through the head-major
    copy of the bias; GQA, ragged lengths and a key padding mask on top."""
    # (QL * QH odd in the last two: the 16-bit layout change then moves single elements instead of pairs)
    for (B, QH, KH, QL, KL) in [(2, 2, 2, 300, 300), (1, 4, 2, 513, 513), (2, 2, 1, 255, 640), (1, 3, 3, 1024, 1024),
                                (1, 3, 1, 255, 255), (2, 3, 3, 129, 300)]:
        if causal and QL != KL:
            continue
        q, k, v, dO, pr, m = _inputs(B, QH, KH, QL, KL, E, dtype, QL + E, pair=True, mask=True)
        _check(nnop, q, k, v, dO, pr, m, causal, H16_TOL, expect_path=1)


def test_pair_without_workspace_falls_to_generic(nnop):
    """The plain forward entry point has no workspace for the head-major copy of the bias: it serves
    `pair` with the SIMT kernels and must agree with the tensor-core result."""
    q, k, v, dO, pr, m = _inputs(2, 4, 2, 384, 384, 64, torch.float32, 9, pair=True, mask=True)
    qd, kd, vd, pd, md = (t.cuda() for t in (q, k, v, pr, m))
    o1, lse1 = nnop._flash_attention(qd, kd, vd, pd, causal=True, kpad_mask=md)
    assert nnop.last_attention_path() == 1
    o0 = torch.empty_like(qd)
    lse0 = torch.empty_like(lse1)
    from nnop_b200._lib import lib, check
    check(lib.nnop_flash_attn_fwd(o0.data_ptr(), lse0.data_ptr(), qd.data_ptr(), kd.data_ptr(), vd.data_ptr(),
                                  pd.data_ptr(), md.data_ptr(), 0, 64, 384, 384, 4, 2, 2, 1, 0.125,
                                  torch.cuda.current_stream().cuda_stream))
    assert nnop.last_attention_path() == 0
    assert max_abs(o0, o1) < F32_TOL and max_abs(lse0, lse1) < F32_TOL


def test_pair_reference_benchmark_shape(nnop):
    """benchmarks/main.jl:305-386: Float32 E=64 L=2048 H=4 B=4 with pair and pad mask, one (b, h) slab
    checked against the oracle (the full problem's score tensor is 268 MB per batch element in fp64)."""
    B, H, L, E = 4, 4, 2048, 64
    q, k, v, dO, pr, m = _inputs(B, H, H, L, L, E, torch.float32, 77, pair=True, mask=True)
    dev = lambda t: t.cuda()
    for causal in (False, True):
        o, lse = nnop._flash_attention(dev(q), dev(k), dev(v), dev(pr), causal=causal, kpad_mask=dev(m))
        assert nnop.last_attention_path() == 1
        dq, dk, dv, dpair = nnop.grad_flash_attention(dev(dO), o, lse, dev(q), dev(k), dev(v), dev(pr),
                                                      causal=causal, kpad_mask=dev(m))
        assert nnop.last_attention_path() == 1
        b, h = B - 1, 2   # the batch element that carries the padding
        sl = lambda t: t[b:b + 1, h:h + 1].double()
        prs = pr[b:b + 1, :, :, h:h + 1].double()
        ro, rl = O.naive_attention(sl(q), sl(k), sl(v), prs, causal=causal, kpad_mask=m[b:b + 1], return_lse=True)
        rq, rk, rv, rp = O.naive_attention_bwd(sl(dO), sl(q), sl(k), sl(v), prs, causal=causal, kpad_mask=m[b:b + 1])
        assert max_abs(o[b:b + 1, h:h + 1], ro) < F32_TOL and max_abs(lse[b:b + 1, h:h + 1], rl) < F32_TOL
        assert max_abs(dq[b:b + 1, h:h + 1], rq) < F32_TOL
        assert max_abs(dk[b:b + 1, h:h + 1], rk) < F32_TOL and max_abs(dv[b:b + 1, h:h + 1], rv) < F32_TOL
        assert max_abs(dpair[b:b + 1, :, :, h:h + 1], rp) < F32_TOL


def _zero_masked_check(nnop, q, k, v, dO, m, causal):
    """kpad_mask with fully masked rows: outputs 0 / lse -inf instead of the reference's NaN."""
    dev = lambda t: t.cuda()
    o, lse = nnop._flash_attention(dev(q), dev(k), dev(v), causal=causal, kpad_mask=dev(m))
    assert nnop.last_attention_path() == 1
    D = lambda t: t.double()
    ro, rl = O.naive_attention(D(q), D(k), D(v), causal=causal, kpad_mask=m, zero_masked_rows=True, return_lse=True)
    assert max_abs(o, ro) < H16_TOL
    fin = torch.isfinite(rl)
    assert torch.equal(torch.isfinite(lse.cpu()), fin) and max_abs(lse.cpu()[fin], rl[fin]) < 1e-3
    dq, dk, dv, _ = nnop.grad_flash_attention(dev(dO), o, lse, dev(q), dev(k), dev(v), causal=causal, kpad_mask=dev(m))
    rq, rk, rv, _ = O.naive_attention_bwd(D(dO), D(q), D(k), D(v), causal=causal, kpad_mask=m, zero_masked_rows=True)
    assert kernel_err(dq, rq) < H16_TOL and kernel_err(dk, rk) < H16_TOL and kernel_err(dv, rv) < H16_TOL


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("E", [128, 64, 256])   # 256: the one-q-tile-per-CTA forward (its backward is SIMT)
@pytest.mark.parametrize("causal", [False, True])
def test_tcgen05_kpad_mask(nnop, dtype, E, causal):
    """kpad_mask on the tensor-core path: the reference's own pattern (test/attention_tests.jl:27-28,
    last 11 keys of the last batch element), prefix-shaped padding (trailing blocks never loaded),
    scattered holes, and an entirely padded batch element."""
    for (B, QH, KH, QL, KL) in [(3, 2, 2, 255, 255), (2, 4, 2, 512, 512), (2, 2, 1, 300, 700)]:
        if causal and QL != KL:
            continue
        q, k, v, dO, _, m = _inputs(B, QH, KH, QL, KL, E, dtype, QL + KL, mask=True)
        _check(nnop, q, k, v, dO, None, m, causal, H16_TOL, expect_path=1)
        # prefix-shaped: batch element b keeps only its first len_b keys
        mp = torch.zeros(B, KL, dtype=torch.bool)
        for b, n in enumerate([KL, KL // 3 + 1, 129][:B]):
            mp[b, :n] = True
        if not causal:   # with a causal mask every row still sees key 0, so no row is empty
            _check(nnop, q, k, v, dO, None, mp, causal, H16_TOL, expect_path=1)
        else:
            _check(nnop, q, k, v, dO, None, mp, causal, H16_TOL, expect_path=1)
        # scattered holes (key 0 kept so that no row is empty under the causal mask)
        g = torch.Generator().manual_seed(KL)
        ms = torch.rand(B, KL, generator=g) > 0.3
        ms[:, 0] = True
        _check(nnop, q, k, v, dO, None, ms, causal, H16_TOL, expect_path=1)
        # one batch element entirely padded, and a hole at key 0 (causal row 0 then sees nothing)
        mz = ms.clone()
        mz[-1, :] = False
        mz[0, 0] = False
        _zero_masked_check(nnop, q, k, v, dO, mz, causal)


def test_full_size_properties_config_c2(nnop):
    """BASELINE config C2 (bf16 causal E=128 L=8192 H=32 B=8): properties that need no oracle run.
    (1) V = const rows => O = that constant; (2) row 0 of a causal O is v[0]; (3) lse of q = 0 is
    log(#visible keys); (4) sum_k dV[k] = sum_q dO[q] (softmax rows sum to one); (5) backward is linear
    in dO; (6) one (b,h) slab agrees with the oracle."""
    B, H, L, E = 8, 32, 8192, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, H, L, E, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
    k = torch.randn(B, H, L, E, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
    v = torch.randn(B, H, L, E, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
    o, lse = nnop._flash_attention(q, k, v, causal=True)
    assert nnop.last_attention_path() == 1
    assert torch.isfinite(o).all() and torch.isfinite(lse).all()
    assert (o[:, :, 0].float() - v[:, :, 0].float()).abs().max().item() < 1e-2          # (2)
    vc = torch.full_like(v, 0.5)
    oc, _ = nnop._flash_attention(q, k, vc, causal=True)
    assert (oc.float() - 0.5).abs().max().item() < 4e-3                                  # (1)
    del oc, vc
    _, lse0 = nnop._flash_attention(torch.zeros_like(q[:1]), k[:1], v[:1], causal=True)
    ref = torch.log(torch.arange(1, L + 1, device="cuda", dtype=torch.float32))
    assert (lse0 - ref).abs().max().item() < 1e-4                                        # (3)
    dO = torch.randn(B, H, L, E, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
    dq, dk, dv, _ = nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
    assert torch.isfinite(dq).all() and torch.isfinite(dk).all() and torch.isfinite(dv).all()
    lhs = dv.float().sum(dim=2)
    rhs = dO.float().sum(dim=2)
    assert (lhs - rhs).abs().max().item() < 2e-2 * math.sqrt(L)                          # (4)
    dq2, dk2, dv2, _ = nnop.grad_flash_attention((dO.float() * 2).to(torch.bfloat16), o, lse, q, k, v, causal=True)
    assert (dq2.float() - 2 * dq.float()).abs().max().item() < 5e-2                       # (5)
    assert (dv2.float() - 2 * dv.float()).abs().max().item() < 5e-2
    del dq2, dk2, dv2
    b, h = 5, 17                                                                          # (6)
    sl = lambda t: t[b:b + 1, h:h + 1].cpu().double()
    o_ref, lse_ref = O.naive_attention(sl(q), sl(k), sl(v), causal=True, return_lse=True)
    assert max_abs(o[b:b + 1, h:h + 1], o_ref) < H16_TOL and max_abs(lse[b:b + 1, h:h + 1], lse_ref) < 1e-3
    rq, rk, rv, _ = O.naive_attention_bwd(sl(dO), sl(q), sl(k), sl(v), causal=True)
    assert kernel_err(dq[b:b + 1, h:h + 1], rq) < H16_TOL
    assert kernel_err(dk[b:b + 1, h:h + 1], rk) < H16_TOL
    assert kernel_err(dv[b:b + 1, h:h + 1], rv) < H16_TOL


@pytest.mark.parametrize("B,QH,KH", [(8, 32, 32), (2, 32, 8)], ids=["C2", "C3-gqa"])
def test_config_c2_every_slab_against_torch_sdpa(nnop, B, QH, KH):
    """BASELINE config C2 (and C3: GQA 32 / 8, two batch elements) at its full shape, EVERY (b, h) slab, O and all three gradients, against an independent
    GPU implementation (torch's scaled_dot_product_attention + autograd, one batch element at a time).  The oracle
    stays the judge of accuracy (one slab above, the small grids elsewhere); this catches what a single slab cannot:
    a tile or head that goes wrong only at some (b, h, block) of the full grid.  Both sides round to bf16 with
    different summation orders, so the bound is BASELINE's 2e-2 plus two bf16 ulps of the element."""
    import torch.nn.functional as F
    L, E = 8192, 128
    g = torch.Generator(device="cuda").manual_seed(1)
    rnd = lambda h: torch.randn(B, h, L, E, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
    q, k, v, dO = rnd(QH), rnd(KH), rnd(KH), rnd(QH)
    o, lse = nnop._flash_attention(q, k, v, causal=True)
    assert nnop.last_attention_path() == 1
    dq, dk, dv, _ = nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
    assert nnop.last_attention_path() == 1

    def close(got, ref, what, b):
        d = (got.float() - ref.float()).abs()
        bound = H16_TOL + 2 * ulp_T(ref, torch.bfloat16).float()
        bad = d > bound
        assert not bad.any(), f"{what}, batch element {b}: {int(bad.sum())} elements off, worst {d.max().item():.4f}"
    for b in range(B):
        qb, kb, vb = (t[b:b + 1].clone().requires_grad_(True) for t in (q, k, v))
        ob = F.scaled_dot_product_attention(qb, kb, vb, is_causal=True, enable_gqa=QH != KH)
        gq, gk, gv = torch.autograd.grad(ob, (qb, kb, vb), dO[b:b + 1])
        close(o[b:b + 1], ob.detach(), "o", b)
        close(dq[b:b + 1], gq, "dq", b)
        close(dk[b:b + 1], gk, "dk", b)
        close(dv[b:b + 1], gv, "dv", b)


@pytest.mark.parametrize("causal", [False, True])
def test_pair_backward_matches_single_cta(nnop, causal):
    """The experimental CTA-pair backward (tcgen05 cta_group::2) must reproduce the default kernel:
    dK / dV bit for bit, dQ up to the order of its fp32 atomics.  Fresh tensors on every trial:
    cold pages once exposed a race between the two compute warpgroups on the P^T / S^T alias."""
    try:
        for trial, (B, QH, KH, QL, KL) in enumerate([(1, 1, 1, 512, 256), (2, 4, 2, 1024, 1024), (1, 2, 2, 300, 700),
                                                     (1, 2, 1, 1000, 1000), (2, 2, 2, 640, 640), (1, 1, 1, 384, 256)] * 3):
            if causal and QL != KL:
                continue
            q, k, v, dO, _, _ = _inputs(B, QH, KH, QL, KL, 128, torch.bfloat16, trial)
            qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
            o, lse = nnop._flash_attention(qd, kd, vd, causal=causal)
            nnop.set_bwd_pair_mode(0)
            ref = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
            nnop.set_bwd_pair_mode(1)
            got = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
            assert max_abs(got[0], ref[0]) <= 2 ** -7 * max(1.0, ref[0].abs().max().item())
            assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2])
    finally:
        nnop.set_bwd_pair_mode(0)


@pytest.mark.parametrize("causal", [False, True])
def test_bitwise_repeatability_on_cold_buffers(nnop, causal):
    """O, lse, dK and dV have a fixed summation order, so repeated runs must agree bit for bit -- also
    when every run touches freshly allocated (cold) memory, which shifts the timing between the
    warp-specialised roles.  (dQ is an fp32 atomic reduction: equal up to its last bits.)"""
    B, QH, KH, L, E = 2, 4, 2, 1024, 128
    q, k, v, dO, _, _ = _inputs(B, QH, KH, L, L, E, torch.bfloat16, 21)
    keep, first = [], None
    for trial in range(12):
        qd, kd, vd, dOd = (x.clone().cuda() for x in (q, k, v, dO))   # new device buffers every time
        keep.append((qd, kd, vd, dOd))
        o, lse = nnop._flash_attention(qd, kd, vd, causal=causal)
        dq, dk, dv, _ = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
        if first is None:
            first = (o, lse, dq, dk, dv)
            continue
        assert torch.equal(o, first[0]) and torch.equal(lse, first[1]), f"forward differs on run {trial}"
        assert torch.equal(dk, first[3]) and torch.equal(dv, first[4]), f"dK/dV differ on run {trial}"
        assert max_abs(dq, first[2]) <= 2 ** -7 * max(1.0, first[2].abs().max().item())


@pytest.mark.parametrize("causal", [False, True])
def test_f32_tensor_core_forward_vs_simt_and_oracle(nnop, causal):
    """Float32 E = 64 forward on the tensor cores (operands split into two bf16 terms, fp32
    accumulation; README config C1 family): within BASELINE.json's 1e-4 of the fp64 oracle, and
    close to the SIMT fp32 path, on the reference's ragged lengths, with GQA and a key padding mask."""
    for (B, QH, KH, QL, KL) in [(2, 4, 4, 1024, 1024), (1, 4, 2, 255, 511), (2, 6, 2, 513, 513), (1, 2, 2, 4096, 4096)]:
        if causal and QL != KL:
            continue
        q, k, v, _, _, m = _inputs(B, QH, KH, QL, KL, 64, torch.float32, QL + 3, mask=True)
        qd, kd, vd, md = q.cuda(), k.cuda(), v.cuda(), m.cuda()
        o, lse = nnop._flash_attention(qd, kd, vd, causal=causal, kpad_mask=md)
        assert nnop.last_attention_path() == 1
        ro, rl = O.naive_attention(q.double(), k.double(), v.double(), causal=causal, kpad_mask=m, return_lse=True)
        assert max_abs(o, ro) < F32_TOL and max_abs(lse, rl) < F32_TOL
        try:
            nnop.set_attention_path(1)
            o_s, lse_s = nnop._flash_attention(qd, kd, vd, causal=causal, kpad_mask=md)
            assert nnop.last_attention_path() == 0
        finally:
            nnop.set_attention_path(0)
        assert max_abs(o, o_s) < F32_TOL and max_abs(lse, lse_s) < F32_TOL


@pytest.mark.parametrize("chunk,kv_heads", [(1, 0), (2, 0), (1, 1), (1, 2), (1, 3)])
def test_host_pipeline_matches_device_call(nnop, chunk, kv_heads):
    """HostAttentionPipeline (pinned host arrays, chunked over (kv-head group, batch)) returns bit for
    bit what the device-resident call returns: every unit is independent (src/attention.jl:152)."""
    B, QH, KH, L, E = 3, 8, 4, 384, 128
    q, k, v, dO, _, _ = _inputs(B, QH, KH, L, L, E, torch.bfloat16, seed=7)
    dq_, dk_, dv_, ddO = (t.cuda() for t in (q, k, v, dO))
    o, lse = nnop._flash_attention(dq_, dk_, dv_, causal=True)
    gq, gk, gv, _ = nnop.grad_flash_attention(ddO, o, lse, dq_, dk_, dv_, causal=True)
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
    out = {n: torch.zeros(s.shape, dtype=s.dtype).pin_memory() for n, s in
           (("o", q), ("dq", q), ("dk", k), ("dv", k))}
    pipe = nnop.HostAttentionPipeline(q.shape, k.shape, torch.bfloat16, causal=True, chunk=chunk,
                                      kv_heads=kv_heads)
    for _ in range(2):   # second call reuses the staging slots
        pipe(pin(q), pin(k), pin(v), pin(dO), out)
    for name, ref in (("o", o), ("dq", gq), ("dk", gk), ("dv", gv)):
        assert torch.equal(out[name], ref.cpu()), name
    nbytes = lambda *ts: sum(t.numel() * t.element_size() for t in ts)
    assert pipe.h2d_bytes == nbytes(q, k, v, dO) and pipe.d2h_bytes == nbytes(q, q, k, k)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("E", [64, 128])
def test_persistent_backward_matches_one_cta_per_tile(nnop, causal, E):
    """The persistent backward (dynamic tile queue, dK / dV epilogue overlapped with the next tile) keeps
    the per-tile summation order of the one-CTA-per-tile kernel: dK / dV bit for bit, dQ up to the
    order of its fp32 reduce-adds.  Run on all SMs (mode 3) and squeezed onto 1 and 3 CTAs (modes
    101, 103) so that every CTA walks many tiles of different lengths back to back."""
    try:
        for trial, (B, QH, KH, QL, KL) in enumerate([(1, 1, 1, 512, 256), (2, 4, 2, 1024, 1024), (1, 2, 2, 300, 700),
                                                     (1, 2, 1, 1000, 1000), (3, 2, 2, 640, 640), (1, 1, 1, 128, 128),
                                                     (2, 3, 3, 129, 129), (1, 6, 2, 2048, 2048)]):
            if causal and QL != KL:
                continue
            dtype = torch.bfloat16 if trial % 2 == 0 else torch.float16
            q, k, v, dO, _, _ = _inputs(B, QH, KH, QL, KL, E, dtype, 100 + trial)
            qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
            o, lse = nnop._flash_attention(qd, kd, vd, causal=causal)
            nnop.set_bwd_pair_mode(2)
            ref = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
            for mode in (3, 101, 103, 3):
                nnop.set_bwd_pair_mode(mode)
                got = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
                assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]), (mode, B, QH, KH, QL, KL)
                assert max_abs(got[0], ref[0]) <= 2 ** -7 * max(1.0, ref[0].abs().max().item())
    finally:
        nnop.set_bwd_pair_mode(0)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("E", [64, 128])
def test_persistent_forward_matches_one_cta_per_tile(nnop, causal, E):
    """The persistent forward (dynamic queue of (q tile, head, batch) work tiles, next tile's Q / K / V
    loaded under the current one, O stored through private staging) runs the same softmax code as the
    one-CTA-per-tile kernel: O and lse bit for bit.  On all SMs (mode 2) and squeezed onto 1 / 3 CTAs
    (modes 101 / 103) so that every CTA walks many tiles of different lengths back to back."""
    try:
        for trial, (B, QH, KH, QL, KL) in enumerate([(1, 1, 1, 512, 256), (2, 4, 2, 1024, 1024), (1, 2, 2, 300, 700),
                                                     (1, 2, 1, 1000, 1000), (3, 2, 2, 640, 640), (1, 1, 1, 128, 128),
                                                     (2, 3, 3, 129, 129), (1, 6, 2, 2048, 2048), (2, 2, 2, 257, 257),
                                                     (1, 1, 1, 1, 1), (1, 2, 2, 255, 64)]):
            if causal and QL != KL:
                continue
            dtype = torch.bfloat16 if trial % 2 == 0 else torch.float16
            q, k, v, dO, _, _ = _inputs(B, QH, KH, QL, KL, E, dtype, 200 + trial)
            qd, kd, vd = q.cuda(), k.cuda(), v.cuda()
            nnop.set_fwd_mode(1)
            o_ref, lse_ref = nnop._flash_attention(qd, kd, vd, causal=causal)
            for mode in (2, 101, 103, 2):
                nnop.set_fwd_mode(mode)
                o, lse = nnop._flash_attention(qd, kd, vd, causal=causal)
                assert torch.equal(o, o_ref) and torch.equal(lse, lse_ref), (mode, B, QH, KH, QL, KL)
    finally:
        nnop.set_fwd_mode(0)


@pytest.mark.parametrize("causal", [False, True])
def test_quad_forward_matches_default_kernel(nnop, causal):
    """nnop_set_fwd_mode(3), the experiment with two softmax warps per 32 query rows (half a row per thread,
    partial row max / row sum exchanged through shared memory): same P bit for bit, only the
    row sum is added in a different order -- O within one ulp of T, lse within fp32 rounding -- incl. ragged
    sizes, GQA and a key padding mask."""
    try:
        for trial, (B, QH, KH, QL, KL, mask) in enumerate([(2, 4, 4, 1024, 1024, False), (1, 4, 2, 300, 517, False),
                                                           (3, 2, 2, 255, 1024, True), (1, 2, 1, 129, 64, False),
                                                           (2, 2, 2, 1000, 1000, True), (1, 1, 1, 1, 1, False)]):
            if causal and QL != KL:
                continue
            dtype = torch.bfloat16 if trial % 2 == 0 else torch.float16
            q, k, v, _, _, m = _inputs(B, QH, KH, QL, KL, 128, dtype, 700 + trial, mask=mask)
            qd, kd, vd = q.cuda(), k.cuda(), v.cuda()
            md = m.cuda() if m is not None else None
            nnop.set_fwd_mode(1)
            o_ref, lse_ref = nnop._flash_attention(qd, kd, vd, causal=causal, kpad_mask=md)
            nnop.set_fwd_mode(3)
            o, lse = nnop._flash_attention(qd, kd, vd, causal=causal, kpad_mask=md)
            assert nnop.last_attention_path() == 1
            ulp = 2.0 ** (-8 if dtype == torch.bfloat16 else -11)
            assert max_abs(o, o_ref) <= ulp * max(1.0, o_ref.float().abs().max().item()), (B, QH, KH, QL, KL)
            fin = torch.isfinite(lse_ref)
            assert torch.equal(torch.isfinite(lse), fin)
            assert max_abs(lse[fin], lse_ref[fin]) <= 1e-5 * max(1.0, lse_ref[fin].abs().max().item())
    finally:
        nnop.set_fwd_mode(0)


@pytest.mark.parametrize("dtype,E", [(torch.float32, 64), (torch.bfloat16, 128)])
def test_backward_reuses_forward_pair_copy(nnop, dtype, E):
    """The head-major copy of `pair` the forward leaves in its workspace can be handed to the backward
    (nnop_flash_attn_bwd_reuse_pair; the rrule keeps it as its third residual, where the reference
    keeps `ls`): same gradients bit for bit (dQ up to its atomics), one layout change less."""
    q, k, v, dO, pr, m = _inputs(2, 4, 2, 384, 384, E, dtype, 31, pair=True, mask=True)
    qd, kd, vd, dOd, pd, md = (t.cuda() for t in (q, k, v, dO, pr, m))
    o, lse, hm = nnop._flash_attention(qd, kd, vd, pd, causal=True, kpad_mask=md, keep_pair_copy=True)
    assert hm is not None and nnop.last_attention_path() == 1
    ref = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, pd, causal=True, kpad_mask=md)
    got = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, pd, causal=True, kpad_mask=md, pair_head_major=hm)
    assert nnop.last_attention_path() == 1
    assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]) and torch.equal(got[3], ref[3])
    assert max_abs(got[0], ref[0]) <= 2 ** -7 * max(1.0, ref[0].abs().max().item())
    # and through autograd (the public flash_attention)
    qa, ka, va, pa = (t.clone().requires_grad_(True) for t in (qd, kd, vd, pd))
    nnop.flash_attention(qa, ka, va, pa, causal=True, kpad_mask=md).backward(dOd)
    assert torch.equal(ka.grad, ref[1]) and torch.equal(va.grad, ref[2]) and torch.equal(pa.grad, ref[3])


@pytest.mark.parametrize("E,L", [(64, 1024), (128, 2048), (64, 4096)])
def test_automatic_kernel_choice_is_transparent(nnop, E, L):
    """Shapes for which the automatic choice takes the persistent forward (E = 64 or QL <= 2048, tile queue
    at least two rounds deep) and the persistent backward: results must not depend on the choice --
    O, lse, dK, dV bit for bit against the one-CTA-per-tile kernels, dQ up to its fp32 reduce order."""
    B, H = 8, 16
    q, k, v, dO, _, _ = _inputs(B, H, H, L, L, E, torch.bfloat16, 300 + E)
    qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
    try:
        nnop.set_fwd_mode(1)
        nnop.set_bwd_pair_mode(2)
        o_ref, lse_ref = nnop._flash_attention(qd, kd, vd, causal=True)
        ref = nnop.grad_flash_attention(dOd, o_ref, lse_ref, qd, kd, vd, causal=True)
        nnop.set_fwd_mode(0)
        nnop.set_bwd_pair_mode(0)
        o, lse = nnop._flash_attention(qd, kd, vd, causal=True)
        got = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=True)
    finally:
        nnop.set_fwd_mode(0)
        nnop.set_bwd_pair_mode(0)
    assert torch.equal(o, o_ref) and torch.equal(lse, lse_ref)
    assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2])
    assert max_abs(got[0], ref[0]) <= 2 ** -7 * max(1.0, ref[0].abs().max().item())


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("E", [64, 128])
def test_duo_backward_matches_one_cta_per_tile(nnop, causal, E):
    """Backward mode 4: persistent CTA pairs (clusters of two) that own kv blocks (2jp, 2jp+1), walk the same
    q blocks and exchange halves of their dQ_i partials over distributed shared memory, so that each CTA
    issues one L2 reduce-add of half the width.  dK / dV keep the per-tile summation order (CTA 1's extra,
    fully masked first causal step adds exact zeros): bit for bit; dQ up to the order of its fp32 adds.
    Shapes cover an odd number of kv blocks (the last pair's second block is out of range), ragged lengths,
    QL != KL, GQA and many tiles per cluster."""
    try:
        for trial, (B, QH, KH, QL, KL) in enumerate([(1, 1, 1, 512, 256), (2, 4, 2, 1024, 1024), (1, 2, 2, 300, 700),
                                                     (1, 2, 1, 1000, 1000), (3, 2, 2, 640, 640), (1, 1, 1, 384, 384),
                                                     (2, 3, 3, 257, 257), (1, 6, 2, 2048, 2048), (4, 8, 8, 1536, 1536)]):
            if causal and QL != KL:
                continue
            dtype = torch.bfloat16 if trial % 2 == 0 else torch.float16
            q, k, v, dO, _, _ = _inputs(B, QH, KH, QL, KL, E, dtype, 400 + trial)
            qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
            o, lse = nnop._flash_attention(qd, kd, vd, causal=causal)
            nnop.set_bwd_pair_mode(2)
            ref = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
            nnop.set_bwd_pair_mode(4)
            for rep in range(2):
                got = nnop.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=causal)
                assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]), (rep, B, QH, KH, QL, KL)
                assert max_abs(got[0], ref[0]) <= 2 ** -7 * max(1.0, ref[0].abs().max().item()), (B, QH, KH, QL, KL)
    finally:
        nnop.set_bwd_pair_mode(0)


def test_stateless_abi_graph_replay_and_threads(nnop):
    """The product ABI keeps no device-side or process-wide state (round 2: the persistent kernels' tile counters
    live in the caller's workspace).  (1) A forward + backward captured into a CUDA graph replays to the same
    bits, also when two captured graphs with their own workspaces replay concurrently on two streams.  (2) Two
    host threads driving different streams at the same time get the results a single thread gets."""
    import threading
    B, H, L, E = 8, 16, 1024, 64          # E = 64, deep tile queue: persistent forward and backward
    mk = lambda seed: tuple(t.cuda() for t in _inputs(B, H, H, L, L, E, torch.bfloat16, seed)[:4])
    sets = [mk(900), mk(901)]
    ref = []
    for q, k, v, dO in sets:
        o, lse = nnop._flash_attention(q, k, v, causal=True)
        ref.append((o, lse) + tuple(nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)[:3]))
    torch.cuda.synchronize()

    # (1) graphs
    graphs, outs, streams = [], [], [torch.cuda.Stream(), torch.cuda.Stream()]
    for (q, k, v, dO), st in zip(sets, streams):
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            o, lse = nnop._flash_attention(q, k, v, causal=True)          # warm-up on the capture stream
            nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            o, lse = nnop._flash_attention(q, k, v, causal=True)
            dq, dk, dv, _ = nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
        graphs.append(g)
        outs.append((o, lse, dq, dk, dv))
    for _ in range(3):
        for g, st in zip(graphs, streams):       # both graphs in flight at once, each on its own stream
            with torch.cuda.stream(st):
                g.replay()
    torch.cuda.synchronize()
    for got, want in zip(outs, ref):
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert torch.equal(got[3], want[3]) and torch.equal(got[4], want[4])
        assert max_abs(got[2], want[2]) <= 2 ** -7 * max(1.0, want[2].abs().max().item())

    # (2) threads
    results, errors = [None, None], []

    def work(i):
        try:
            st = torch.cuda.Stream()
            q, k, v, dO = sets[i]
            with torch.cuda.stream(st):
                for _ in range(4):
                    o, lse = nnop._flash_attention(q, k, v, causal=True)
                    dq, dk, dv, _ = nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
            st.synchronize()
            results[i] = (o, lse, dq, dk, dv)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for got, want in zip(results, ref):
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert torch.equal(got[3], want[3]) and torch.equal(got[4], want[4])
