"""online_softmax / rms_norm / layer_norm through the C ABI vs the oracle, on the reference's
own grids (test/softmax_tests.jl:12-29, test/rmsnorm_tests.jl:11-33, test/layernorm_tests.jl:13-35),
plus golden vectors, 16-bit types and the large shapes of BASELINE config C3."""
import pytest
import torch

from helpers import load_golden, max_abs, reference_isapprox
from oracle import oracle as O

pytestmark = pytest.mark.gpu

EMBS = (15, 255, 256, 257, 511, 512, 513, 1024)
NS = (1, 2, 4, 15, 16, 17, 23, 25)
# element-wise bound for fp32 (inputs are U[0,1), outputs O(1)); the reference asserts 1f-6 norm-wise
F32_TOL = 2e-5


def _ones_grad(y):
    return torch.ones_like(y)


@pytest.mark.parametrize("seq_len", [32, 33, 63, 255, 256, 511, 512, 513, 1024, 8192, 20000])
def test_softmax_reference_grid(nnop, seq_len):
    g = torch.Generator().manual_seed(seq_len)
    x = torch.rand(4, seq_len, generator=g)
    dy = torch.randn(4, seq_len, generator=g)
    xd = x.cuda().requires_grad_(True)
    y = nnop.online_softmax(xd)
    y_ref = O.naive_softmax(x.double())
    assert max_abs(y, y_ref) < 1e-6
    assert reference_isapprox(y, y_ref, 0.0, 3.5e-4)      # `y1 ≈ y2`
    (gx,) = torch.autograd.grad(y, xd, dy.cuda())
    assert max_abs(gx, O.naive_softmax_bwd(dy.double(), y_ref)) < 1e-6
    (g1,) = torch.autograd.grad(nnop.online_softmax(xd).sum(), xd)   # the reference's sum-gradient
    assert reference_isapprox(g1, torch.zeros_like(x), 1e-6, 1e-6)


@pytest.mark.parametrize("emb", EMBS)
@pytest.mark.parametrize("offset", [0.0, 1.0])
def test_rms_norm_reference_grid(nnop, emb, offset):
    for n in NS:
        g = torch.Generator().manual_seed(emb * 100 + n)
        x = torch.rand(n, emb, generator=g)
        w = torch.rand(emb, generator=g)
        dy = torch.randn(n, emb, generator=g)
        y, rstd = nnop._rms_norm(x.cuda(), w.cuda(), eps=1e-6, offset=offset)
        y_ref, rstd_ref = O.naive_rms_norm(x.double(), w.double(), offset=offset, return_rstd=True)
        assert max_abs(y, y_ref) < F32_TOL, (emb, n)
        assert max_abs(rstd, rstd_ref) < F32_TOL
        for d in (dy, torch.ones_like(dy)):
            dx, dw = nnop.grad_rms_norm(d.cuda(), rstd, x.cuda(), w.cuda(), offset=offset)
            dx_ref, dw_ref = O.naive_rms_norm_bwd(d.double(), x.double(), w.double(), offset=offset)
            assert dw.dtype == torch.float32
            assert max_abs(dx, dx_ref) < F32_TOL, (emb, n)
            assert max_abs(dw, dw_ref) < F32_TOL * max(1, n) ** 0.5, (emb, n)


@pytest.mark.parametrize("emb", EMBS)
def test_layer_norm_reference_grid(nnop, emb):
    for n in NS:
        g = torch.Generator().manual_seed(emb * 100 + n)
        x = torch.rand(n, emb, generator=g)
        w = torch.rand(emb, generator=g)
        b = torch.rand(emb, generator=g)
        dy = torch.randn(n, emb, generator=g)
        y, mean, rstd = nnop._layer_norm(x.cuda(), w.cuda(), b.cuda(), eps=1e-6)
        y_ref, mu_ref, rs_ref = O.naive_layer_norm(x.double(), w.double(), b.double(), return_stats=True)
        assert max_abs(y, y_ref) < F32_TOL, (emb, n)
        assert max_abs(mean, mu_ref) < F32_TOL and max_abs(rstd, rs_ref) < 1e-4
        for d in (dy, torch.ones_like(dy)):
            dx, dw, db = nnop.grad_layer_norm(d.cuda(), mean, rstd, x.cuda(), w.cuda(), b.cuda())
            dx_ref, dw_ref, db_ref = O.naive_layer_norm_bwd(d.double(), x.double(), w.double(), b.double())
            assert max_abs(dx, dx_ref) < 1e-4, (emb, n)
            assert max_abs(dw, dw_ref) < 1e-4 and max_abs(db, db_ref) < 1e-4, (emb, n)


def test_autograd_wrappers_match_reference_style(nnop):
    """Zygote.gradient(sum ∘ op) style check through the autograd Functions (the rrule mirrors)."""
    g = torch.Generator().manual_seed(0)
    x = torch.rand(23, 513, generator=g)
    w = torch.rand(513, generator=g)
    b = torch.rand(513, generator=g)
    xd, wd, bd = (t.cuda().requires_grad_(True) for t in (x, w, b))
    gx, gw = torch.autograd.grad(nnop.rms_norm(xd, wd, offset=1.0).sum(), (xd, wd))
    rx, rw = O.naive_rms_norm_bwd(torch.ones(23, 513, dtype=torch.float64), x.double(), w.double(), offset=1.0)
    assert reference_isapprox(gx, rx, 1e-6, 1e-6) and reference_isapprox(gw, rw, 1e-6, 1e-6)
    gx, gw, gb = torch.autograd.grad(nnop.layer_norm(xd, wd, bd).sum(), (xd, wd, bd))
    rx, rw, rb = O.naive_layer_norm_bwd(torch.ones(23, 513, dtype=torch.float64), x.double(), w.double(), b.double())
    assert reference_isapprox(gx, rx, 2e-6, 2e-6) and reference_isapprox(gw, rw, 2e-6, 2e-6)
    assert reference_isapprox(gb, rb, 1e-6, 1e-6)


def test_rowwise_golden(nnop):
    for name, d in load_golden("rowwise.npz").items():
        x = d["x"].float().cuda()
        dy = d["dy"].float().cuda()
        if name.startswith("softmax"):
            y = nnop.online_softmax(x)
            assert max_abs(y, d["y"]) < 1e-6
            assert max_abs(nnop.grad_online_softmax(dy, y), d["dx"]) < 1e-6
        elif name.startswith("rms"):
            off = float(d["offset"])
            y, rstd = nnop._rms_norm(x, d["w"].float().cuda(), offset=off)
            dx, dw = nnop.grad_rms_norm(dy, rstd, x, d["w"].float().cuda(), offset=off)
            assert max_abs(y, d["y"]) < F32_TOL and max_abs(dx, d["dx"]) < F32_TOL and max_abs(dw, d["dw"]) < F32_TOL
        else:
            w, b = d["w"].float().cuda(), d["b"].float().cuda()
            y, mean, rstd = nnop._layer_norm(x, w, b)
            dx, dw, db = nnop.grad_layer_norm(dy, mean, rstd, x, w, b)
            assert max_abs(y, d["y"]) < F32_TOL and max_abs(dx, d["dx"]) < 1e-4
            assert max_abs(dw, d["dw"]) < 1e-4 and max_abs(db, d["db"]) < 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("emb,n", [(4096, 257), (1024, 64), (8192, 33), (10240, 9), (264, 5)])
def test_norms_16bit_and_large(nnop, dtype, emb, n):
    """16-bit storage, fp32 math: error bounded by output rounding (2e-2 abs, BASELINE tolerance)."""
    g = torch.Generator().manual_seed(emb + n)
    x = torch.randn(n, emb, generator=g).to(dtype)
    w = torch.rand(emb, generator=g).to(dtype)
    b = torch.rand(emb, generator=g).to(dtype)
    dy = torch.randn(n, emb, generator=g).to(dtype)
    xd, wd, bd, dyd = x.cuda(), w.cuda(), b.cuda(), dy.cuda()
    X, W, Bb, DY = x.double(), w.double(), b.double(), dy.double()
    y, rstd = nnop._rms_norm(xd, wd)
    assert max_abs(y, O.naive_rms_norm(X, W)) < 2e-2
    dx, dw = nnop.grad_rms_norm(dyd, rstd, xd, wd)
    rx, rw = O.naive_rms_norm_bwd(DY, X, W)
    assert max_abs(dx, rx) < 2e-2 and max_abs(dw, rw) < 2e-3 * n ** 0.5 + 1e-3
    y, mean, rstd = nnop._layer_norm(xd, wd, bd)
    assert max_abs(y, O.naive_layer_norm(X, W, Bb)) < 2e-2
    dx, dw, db = nnop.grad_layer_norm(dyd, mean, rstd, xd, wd, bd)
    rx, rw, rb = O.naive_layer_norm_bwd(DY, X, W, Bb)
    assert max_abs(dx, rx) < 2e-2
    assert max_abs(dw, rw) < 2e-2 * max(1.0, rw.abs().max().item())
    assert max_abs(db, rb) < 2e-2 * max(1.0, rb.abs().max().item())
    s = nnop.online_softmax(xd)
    assert max_abs(s, O.naive_softmax(X)) < 2e-3
    assert max_abs(nnop.grad_online_softmax(dyd, s), O.naive_softmax_bwd(DY, s.double().cpu())) < 2e-2


def test_softmax_rows_sum_to_one_at_bench_size(nnop):
    """size-independent property at the README bench shape (8192 x 1024 f32, benchmarks/main.jl:279-300)."""
    x = torch.randn(1024, 8192, device="cuda")
    y = nnop.online_softmax(x)
    assert (y.sum(-1) - 1).abs().max().item() < 1e-5
    assert torch.equal(y.argmax(-1), x.argmax(-1))
    # shift invariance
    assert (nnop.online_softmax(x + 3.0) - y).abs().max().item() < 1e-6


def test_softmax_second_order(nnop):
    """Second-order AD through `online_softmax` (the reference keeps its pullback differentiable when it is
    itself under differentiation: `within_gradient(y)`, src/softmax.jl:70-74): the gradient of a function of
    the first gradient, against torch autograd on the oracle's fp64 softmax."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(7, 257, generator=g)
    w = torch.randn(7, 257, generator=g)
    u = torch.randn(7, 257, generator=g)

    def second(xx, softmax, ww, uu):
        y = softmax(xx)
        (gx,) = torch.autograd.grad((y * ww).sum(), xx, create_graph=True)   # first-order gradient, kept in the graph
        (hx,) = torch.autograd.grad((gx * uu).sum() + (gx * gx).sum(), xx)
        return gx.detach(), hx

    xr = x.double().requires_grad_(True)
    g_ref, h_ref = second(xr, O.naive_softmax, w.double(), u.double())
    xd = x.cuda().requires_grad_(True)
    g_got, h_got = second(xd, nnop.online_softmax, w.cuda(), u.cuda())
    assert max_abs(g_got, g_ref) < 1e-6
    assert max_abs(h_got, h_ref) < 1e-5
