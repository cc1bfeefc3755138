"""The three HBM-bound helpers of the ring variant (csrc/ring_ops.cu) against torch on one GPU, and the
whole ring schedule driven through CudaBackend with world size 1 (these had no recorded
GPU parity because the 2-GPU NCCL test is skipped on a 1-GPU box)."""
import math

import pytest
import torch

from helpers import kernel_err, max_abs
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.float16, torch.bfloat16]


def _backend():
    from nnop_b200.ring import CudaBackend
    return CudaBackend()


@pytest.mark.parametrize("dtype", DTYPES)
def test_attn_merge_vs_torch(dtype):
    be = _backend()
    g = torch.Generator().manual_seed(1)
    rows, E = 1000, 64
    o_a = torch.randn(rows, E, generator=g)
    o_p = torch.randn(rows, E, generator=g).to(dtype)
    l_a = torch.randn(rows, generator=g) * 3
    l_p = torch.randn(rows, generator=g) * 3
    l_a[5] = l_p[5] = -math.inf          # both sides fully masked -> 0 / -inf
    l_a[6] = -math.inf                   # only the accumulator empty -> takes the partial
    l_p[7] = -math.inf                   # only the partial empty -> keeps the accumulator
    # init: plain copy
    acc = torch.full((rows, E), float("nan"), device="cuda")
    lse = torch.full((rows,), float("nan"), device="cuda")
    out = be.merge(acc, lse, o_p.cuda(), l_p.cuda(), True)
    assert torch.equal(acc.cpu(), o_p.float()) and torch.equal(out.cpu(), l_p)
    # fold
    acc = o_a.clone().cuda()
    out = be.merge(acc, l_a.cuda(), o_p.cuda(), l_p.cuda(), False)
    m = torch.maximum(l_a, l_p).double()
    wa, wp = torch.exp(l_a.double() - m), torch.exp(l_p.double() - m)
    ref = (o_a.double() * wa[:, None] + o_p.double() * wp[:, None]) / (wa + wp)[:, None]
    ref_l = m + torch.log(wa + wp)
    ref[5], ref_l[5] = 0.0, -math.inf
    ref[6], ref_l[6] = o_p[6].double(), l_p[6].double()
    ref[7], ref_l[7] = o_a[7].double(), l_a[7].double()
    assert torch.isfinite(acc).all()
    assert max_abs(acc, ref) < 1e-5
    fin = torch.isfinite(ref_l)
    assert torch.equal(torch.isfinite(out.cpu()), fin) and max_abs(out.cpu()[fin], ref_l[fin]) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
def test_accumulate_and_store_rows_vs_torch(dtype):
    be = _backend()
    g = torch.Generator().manual_seed(2)
    n = 8 * 1237 if dtype != torch.float32 else 4 * 1237      # not a multiple of the CTA size
    part = torch.randn(n, generator=g).to(dtype)
    acc0 = torch.randn(n, generator=g)
    acc = torch.full((n,), float("nan"), device="cuda")
    be.accumulate(acc, part.cuda(), True)
    assert torch.equal(acc.cpu(), part.float())
    acc = acc0.clone().cuda()
    be.accumulate(acc, part.cuda(), False)
    assert torch.equal(acc.cpu(), acc0 + part.float())
    # ragged row windows: (slabs, rows, E) fp32 -> rows [off, off + rows) of (slabs, out_rows, E) T
    slabs, rows, out_rows, E = 6, 37, 101, 32
    a = torch.randn(slabs, rows, E, generator=g)
    for off in (0, 13, out_rows - rows):
        out = torch.full((slabs, out_rows, E), 7.0).to(dtype).cuda()
        be.store_rows(out, a.cuda(), off)
        ref = torch.full((slabs, out_rows, E), 7.0).to(dtype)
        ref[:, off:off + rows] = a.to(dtype)
        assert torch.equal(out.cpu(), ref), off


def test_ring_helpers_reject_misaligned_and_bad_windows(nnop):
    from nnop_b200._lib import lib
    st = torch.cuda.current_stream().cuda_stream
    buf = torch.zeros(4096, device="cuda")
    h = torch.zeros(4096, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(64, device="cuda")
    lse2 = torch.zeros(64, device="cuda")
    assert lib.nnop_accumulate_f32(buf.data_ptr() + 4, h.data_ptr(), 2, 64, 0, st) == 6       # NNOP_ERR_ARG
    assert lib.nnop_accumulate_f32(buf.data_ptr(), h.data_ptr() + 2, 2, 64, 0, st) == 6
    assert lib.nnop_accumulate_f32(buf.data_ptr(), h.data_ptr(), 2, 12, 0, st) == 1           # NNOP_ERR_SHAPE
    assert lib.nnop_accumulate_f32(buf.data_ptr(), buf.data_ptr() + 2048, 0, 12, 0, st) == 0  # Float32: n % 4
    assert lib.nnop_attn_merge(buf.data_ptr() + 8, lse.data_ptr(), lse2.data_ptr(), h.data_ptr(), lse.data_ptr(),
                               2, 64, 16, 0, st) == 6
    assert lib.nnop_attn_merge(buf.data_ptr(), lse.data_ptr(), lse.data_ptr(), h.data_ptr(), lse.data_ptr(),
                               2, 64, 16, 0, st) == 6                                          # aliasing
    assert lib.nnop_store_rows_from_f32(h.data_ptr() + 2, buf.data_ptr(), 2, 64, 1, 4, 8, 0, st) == 6
    assert lib.nnop_store_rows_from_f32(h.data_ptr(), buf.data_ptr(), 2, 64, 1, 4, 8, 5, st) == 1
    assert "row window" in nnop.lib.nnop_last_error_string().decode()
    torch.cuda.synchronize()


@pytest.mark.parametrize("causal", [False, True])
def test_ring_world_size_one_on_one_gpu(nnop, tmp_path, causal):
    """World size 1 (gloo group of one process): the schedule degenerates to the local problem, but every
    CudaBackend call -- attention per chunk pair, merge, accumulate, store_rows -- runs on the GPU."""
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29900 + os.getpid() % 500))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        B, QH, KH, L, E = 1, 4, 2, 512, 128
        g = torch.Generator().manual_seed(3)
        q, k, v, dO = (torch.randn(B, H, L, E, generator=g).to(torch.bfloat16) for H in (QH, KH, KH, QH))
        o, res = nnop.ring_attention_forward(q.cuda(), k.cuda(), v.cuda(), causal=causal)
        dq, dk, dv = nnop.ring_attention_backward(dO.cuda(), res, causal=causal)
    finally:
        dist.destroy_process_group()
    D = lambda t: t.double()
    ro = O.naive_attention(D(q), D(k), D(v), causal=causal)
    rq, rk, rv, _ = O.naive_attention_bwd(D(dO), D(q), D(k), D(v), causal=causal)
    assert max_abs(o, ro) < 2e-2
    assert kernel_err(dq, rq) < 2e-2 and kernel_err(dk, rk) < 2e-2 and kernel_err(dv, rv) < 2e-2


def _p2p_case(nnop, W, devices, shape, dtype, causal, tol=2e-2):
    """W ranks on `devices` (repeats allowed): shard the full problem, run the single-process ring through
    nnop_ring_attn_fwd / _bwd, gather, compare with the oracle's attention over the FULL sequence."""
    B, QH, KH, L, E = shape
    g = torch.Generator().manual_seed(L + W)
    q, k, v, dO = (torch.randn(B, H, L, E, generator=g).to(dtype) for H in (QH, KH, KH, QH))
    shard = (lambda t, r: nnop.zigzag_shard(t, r, W)) if causal else (lambda t, r: nnop.contiguous_shard(t, r, W))
    put = lambda t, r: shard(t, r).to(f"cuda:{devices[r]}")
    qs, ks, vs, dOs = ([put(t, r) for r in range(W)] for t in (q, k, v, dO))
    os_, lses = nnop.p2p_ring_attention_forward(qs, ks, vs, causal=causal)
    dqs, dks, dvs = nnop.p2p_ring_attention_backward(dOs, os_, lses, qs, ks, vs, causal=causal)
    for d in set(devices):
        torch.cuda.synchronize(d)
    gather = (lambda ps: nnop.zigzag_unshard([p.cpu() for p in ps])) if causal else \
        (lambda ps: torch.cat([p.cpu() for p in ps], dim=2))
    o, lse, dq, dk, dv = gather(os_), gather([l.unsqueeze(-1) for l in lses]).squeeze(-1), gather(dqs), gather(dks), gather(dvs)
    D = lambda t: t.double()
    ro, rl = O.naive_attention(D(q), D(k), D(v), causal=causal, return_lse=True)
    rq, rk, rv, _ = O.naive_attention_bwd(D(dO), D(q), D(k), D(v), causal=causal, o=D(o) if tol > 1e-3 else None)
    assert max_abs(o, ro) < tol, "o"
    assert max_abs(lse, rl) < 1e-3, "lse"
    assert kernel_err(dq, rq) < tol, "dq"
    assert kernel_err(dk, rk) < tol, "dk"
    assert kernel_err(dv, rv) < tol, "dv"


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("W", [1, 2, 3, 4])
def test_p2p_ring_virtual_ranks_on_one_gpu(nnop, W, causal):
    """The C-ABI ring (nnop_ring_attn_fwd / _bwd) with W ranks that all live on device 0: the whole schedule,
    the double-buffered landing areas, the event graph between compute and copy streams, the strided
    merges / accumulations and the partial-gradient pushes run exactly as on W GPUs (cudaMemcpyPeerAsync
    between two buffers of one device is an ordinary copy), so a 1-GPU box checks parity of the path."""
    _p2p_case(nnop, W, [0] * W, (1, 4, 2, 256 * W, 128), torch.bfloat16, causal)
    _p2p_case(nnop, W, [0] * W, (2, 2, 2, 192 * W, 64), torch.float16, causal)


def test_p2p_ring_config_c5_full_shape_virtual_ranks(nnop):
    """BASELINE config C5 at its full shape (bf16 causal, B = 1, H = 32, L = 131 072, E = 128, 8 ranks in zig-zag
    order) through nnop_ring_attn_fwd / _bwd with the 8 ranks living on ONE GPU, against the dense kernels on the
    unsharded tensors (themselves checked against the oracle at C2 / C3 shape): O, lse and all three gradients.
    A one-GPU box can hold the whole problem (~25 GB); the schedule, landing buffers, merges and gradient pushes
    are the ones the 8-GPU run uses.  Bound: 2e-2 + 2 bf16 ulps (two bf16 results with different summation orders)."""
    from helpers import ulp_T
    W, B, H, L, E = 8, 1, 32, 131072, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    q, k, v, dO = (torch.randn(B, H, L, E, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
                   for _ in range(4))
    o, lse = nnop._flash_attention(q, k, v, causal=True)
    dq, dk, dv, _ = nnop.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
    qs, ks, vs, dOs = ([nnop.zigzag_shard(t, r, W).contiguous() for r in range(W)] for t in (q, k, v, dO))
    os_, lses = nnop.p2p_ring_attention_forward(qs, ks, vs, causal=True)
    dqs, dks, dvs = nnop.p2p_ring_attention_backward(dOs, os_, lses, qs, ks, vs, causal=True)
    torch.cuda.synchronize()
    del qs, ks, vs, dOs
    un = lambda ps: nnop.zigzag_unshard(ps)
    for got, ref, what in ((un(os_), o, "o"), (un(dqs), dq, "dq"), (un(dks), dk, "dk"), (un(dvs), dv, "dv")):
        d = (got.float() - ref.float()).abs()
        bad = d > 2e-2 + 2 * ulp_T(ref, torch.bfloat16).float()
        assert not bad.any(), f"{what}: {int(bad.sum())} elements off, worst {d.max().item():.4f}"
        del d, bad, got
    l2 = un([l.unsqueeze(-1) for l in lses]).squeeze(-1)
    assert (l2 - lse).abs().max().item() < 1e-3


def test_p2p_ring_float32_and_errors(nnop):
    _p2p_case(nnop, 2, [0, 0], (1, 2, 1, 256, 32), torch.float32, True, tol=1e-4)
    q = torch.randn(1, 2, 255, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(nnop.NNopError, match="even local sequence length"):
        nnop.p2p_ring_attention_forward([q, q], [q, q], [q, q], causal=True)
    with pytest.raises(nnop.NNopError, match="equally shaped"):
        nnop.p2p_ring_attention_forward([q, q[:, :, :128]], [q, q], [q, q], causal=False)
    from nnop_b200._lib import lib
    assert lib.nnop_ring_attn_fwd_workspace_bytes(2, 128, 1024, 8, 2, 1, 4, 1) > 0
    assert lib.nnop_ring_attn_bwd_workspace_bytes(2, 128, 1024, 8, 3, 1, 4, 1) == 0    # QH % KH != 0


@pytest.mark.parametrize("causal", [True, False])
def test_p2p_ring_on_all_gpus(nnop, causal):
    """Real peers: one rank per visible GPU (gpurun --gpus N), K / V and gradient partials over NVLink."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    _p2p_case(nnop, n, list(range(n)), (1, 8, 2, 512 * n, 128), torch.bfloat16, causal)
