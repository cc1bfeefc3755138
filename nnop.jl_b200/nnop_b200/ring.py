"""Sequence-sharded ("ring") flash attention across the GPUs of one node (BASELINE config 5).

Additive: the reference has no multi-GPU path.  One process per GPU
(`torch.distributed`); every rank owns a slice of the sequence axis of q, k, v.  K/V blocks travel
round the ring (rank r -> r+1) over NVLink while the rank attends its queries to the block it
holds with the dense kernels (`nnop_flash_attn_fwd` / `_bwd`); partial results are folded with
`nnop_attn_merge` (log-sum-exp weights, fp32 accumulators).  The next block's transfer is posted
before the current block's kernels, so the copy runs under the math.

Causal attention uses the zig-zag layout: the sequence is cut into 2W chunks and rank r owns
chunks (r, 2W-1-r), which makes every ring step the same amount of work on every rank:
  step 0            local causal attention (chunk pairs (0,0) causal, (1,0) full, (1,1) causal)
  block from s < r  both local query chunks attend the sender's FIRST chunk, no mask
  block from s > r  the local SECOND query chunk attends both of the sender's chunks, no mask
Backward runs the same schedule with `∇flash_attention` on each pair, using the final (merged)
o / lse -- P = exp(S - lse) is then already globally normalised -- with dq accumulated locally
and dk / dv travelling with their K/V block in fp32.

Kernels are reached through a small backend object so that the schedule (this file) can be
exercised on CPU under gloo with a stand-in; the default backend is the CUDA library and refuses
CPU tensors -- there is no fallback.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import NNopError, check, lib
from . import ops


# ------------------------------------------------------------------------------------------
# zig-zag layout helpers (host-side index logic only)
# ------------------------------------------------------------------------------------------
def zigzag_chunks(rank: int, world: int):
    """Global chunk ids (of 2*world) owned by `rank`."""
    return rank, 2 * world - 1 - rank


def zigzag_shard(x: torch.Tensor, rank: int, world: int, dim: int = 2) -> torch.Tensor:
    """Local slice of a full-sequence tensor in zig-zag order (chunks r and 2W-1-r concatenated)."""
    L = x.shape[dim]
    if L % (2 * world) != 0:
        raise NNopError(1, f"sequence length `{L}` must be divisible by 2*world = {2 * world}")
    c = L // (2 * world)
    a, b = zigzag_chunks(rank, world)
    return torch.cat([x.narrow(dim, a * c, c), x.narrow(dim, b * c, c)], dim=dim).contiguous()


def zigzag_unshard(parts, dim: int = 2) -> torch.Tensor:
    """Inverse of `zigzag_shard` given every rank's local tensor (rank order)."""
    world = len(parts)
    c = parts[0].shape[dim] // 2
    chunks = [None] * (2 * world)
    for r, p in enumerate(parts):
        a, b = zigzag_chunks(r, world)
        chunks[a] = p.narrow(dim, 0, c)
        chunks[b] = p.narrow(dim, c, c)
    return torch.cat(chunks, dim=dim)


def contiguous_shard(x: torch.Tensor, rank: int, world: int, dim: int = 2) -> torch.Tensor:
    L = x.shape[dim]
    if L % world != 0:
        raise NNopError(1, f"sequence length `{L}` must be divisible by world = {world}")
    c = L // world
    return x.narrow(dim, rank * c, c).contiguous()


def ring_schedule(rank: int, world: int, causal: bool):
    """Per ring step: list of (q_chunk, kv_chunk, causal_flag) pairs this rank computes on the block
    that originated at rank (rank - step) % world.  Chunk ids are local (0 / 1); non-causal uses
    one chunk per rank."""
    steps = []
    for s in range(world):
        src = (rank - s) % world
        if not causal:
            steps.append((src, [(0, 0, False)]))
        elif s == 0:
            steps.append((src, [(0, 0, True), (1, 0, False), (1, 1, True)]))
        elif src < rank:
            steps.append((src, [(0, 0, False), (1, 0, False)]))
        else:
            steps.append((src, [(1, 0, False), (1, 1, False)]))
    return steps


# ------------------------------------------------------------------------------------------
# kernel backend (CUDA library)
# ------------------------------------------------------------------------------------------
class CudaBackend:
    """The product path: every call is one or two launches of libnnop_b200.so."""

    acc_dtype = torch.float32

    @staticmethod
    def _need_cuda(*ts):
        for t in ts:
            if not t.is_cuda:
                raise NNopError(6, "ring attention runs on CUDA tensors only (there is no CPU path)")

    def attn_fwd(self, q, k, v, causal):
        return ops._flash_attention(q, k, v, causal=causal)

    def attn_bwd(self, dO, o, lse, q, k, v, causal):
        return ops.grad_flash_attention(dO, o, lse, q, k, v, causal=causal)[:3]

    def merge(self, o_acc, lse_acc, o_part, lse_part, init):
        """Fold (o_part, lse_part) into (o_acc, lse_acc); returns the new lse tensor."""
        self._need_cuda(o_acc, o_part)
        lse_new = lse_acc if init else torch.empty_like(lse_acc)
        check(lib.nnop_attn_merge(o_acc.data_ptr(), lse_acc.data_ptr(), lse_new.data_ptr(),
                                  o_part.data_ptr(), lse_part.data_ptr(), ops._dt(o_part),
                                  o_part.shape[-1], lse_part.numel(), int(init), ops._stream()))
        return lse_new

    def accumulate(self, acc, part, init):
        self._need_cuda(acc, part)
        check(lib.nnop_accumulate_f32(acc.data_ptr(), part.data_ptr(), ops._dt(part), part.numel(),
                                      int(init), ops._stream()))

    def store_rows(self, out, acc, row_offset):
        """out[..., row_offset : row_offset + rows, :] = acc (cast to out.dtype)."""
        self._need_cuda(out, acc)
        rows, E = acc.shape[-2], acc.shape[-1]
        check(lib.nnop_store_rows_from_f32(out.data_ptr(), acc.data_ptr(), ops._dt(out), E,
                                           acc.numel() // (rows * E), rows, out.shape[-2], row_offset,
                                           ops._stream()))


# ------------------------------------------------------------------------------------------
# ring exchange
# ------------------------------------------------------------------------------------------
_ring_groups = {}


def ring_group(base_group=None):
    """Process group the K/V rotation runs on.  Under NCCL it is a dedicated group whose kernels
    launch on a HIGH-PRIORITY stream: the attention kernels keep every SM busy (one CTA per SM, all
    shared memory), and on a normal-priority stream a send/recv kernel posted under them only gets
    SMs when the attention grid drains (measured: 0.84 ms -> 11 ms per 268 MB hop).  Collective:
    every rank of `base_group` must make the first call.  Other backends: `base_group` itself."""
    if dist.get_backend(base_group) != "nccl":
        return base_group
    key = id(base_group) if base_group is not None else None
    if key not in _ring_groups:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        ranks = dist.get_process_group_ranks(base_group if base_group is not None else dist.group.WORLD)
        _ring_groups[key] = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
    return _ring_groups[key]


class _Ring:
    def __init__(self, group):
        group = ring_group(group)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.next = dist.get_global_rank(group, (self.rank + 1) % self.world) if group is not None \
            else (self.rank + 1) % self.world
        self.prev = dist.get_global_rank(group, (self.rank - 1) % self.world) if group is not None \
            else (self.rank - 1) % self.world

    def start(self, send: torch.Tensor, recv: torch.Tensor):
        """Post send -> next, recv <- prev; returns handles to wait on."""
        if self.world == 1:
            return []
        ops_ = [dist.P2POp(dist.isend, send, self.next, self.group),
                dist.P2POp(dist.irecv, recv, self.prev, self.group)]
        return dist.batch_isend_irecv(ops_)

    @staticmethod
    def wait(handles):
        for h in handles:
            h.wait()


def _split(x, causal):
    """Chunk-contiguous copies of a local tensor along L: [first half, second half] or [x]."""
    if not causal:
        return [x.contiguous()]
    L = x.shape[2]
    if L % 2 != 0:
        raise NNopError(1, f"causal ring attention needs an even local sequence length, got `{L}`")
    return [x[:, :, :L // 2].contiguous(), x[:, :, L // 2:].contiguous()]


def ring_attention_forward(q, k, v, *, causal: bool, group=None, backend=None):
    """Forward over the local shards q (B,QH,Ll,E), k/v (B,KH,Ll,E) [Julia (E,Ll,H,B)].
    Returns ``(o, residuals)``; residuals feed `ring_attention_backward`."""
    be = backend or CudaBackend()
    ring = _Ring(group)
    qs, ks, vs = _split(q, causal), _split(k, causal), _split(v, causal)
    nch = len(qs)
    kv = torch.stack(ks + vs)            # (2*nch, B, KH, c, E): one buffer per hop
    kv_next = torch.empty_like(kv)
    o_acc = [torch.empty(x.shape, dtype=be.acc_dtype, device=x.device) for x in qs]
    lse = [torch.empty(x.shape[:3], dtype=torch.float32, device=x.device) for x in qs]
    started = [False] * nch
    for s, (src, pairs) in enumerate(ring_schedule(ring.rank, ring.world, causal)):
        handles = ring.start(kv, kv_next) if s + 1 < ring.world else []
        for qc, kc, cz in pairs:
            o_p, lse_p = be.attn_fwd(qs[qc], kv[kc], kv[nch + kc], cz)
            lse[qc] = be.merge(o_acc[qc], lse[qc], o_p, lse_p, not started[qc])
            started[qc] = True
        ring.wait(handles)
        if handles:
            kv, kv_next = kv_next, kv
    o = torch.empty_like(q)
    o_ch = []
    c = qs[0].shape[2]
    for i in range(nch):
        be.store_rows(o, o_acc[i], i * c)
        oc = torch.empty_like(qs[i])
        be.store_rows(oc, o_acc[i], 0)
        o_ch.append(oc)
    return o, (qs, ks, vs, o_ch, lse)


def ring_attention_backward(dO, residuals, *, causal: bool, group=None, backend=None):
    """Backward; returns local ``(dq, dk, dv)`` in the layout of the forward's inputs.

    Both transfers of a step run under its kernels: the next K/V block is prefetched, and the
    travelling dK/dV of the block being worked on (accumulated by the previous rank during the
    previous step) arrives while this rank computes its own contribution into a step-local fp32
    buffer; the two are added afterwards and sent on at the start of the next step."""
    be = backend or CudaBackend()
    ring = _Ring(group)
    qs, ks, vs, o_ch, lse = residuals
    nch = len(qs)
    dOs = _split(dO, causal)
    kv = torch.stack(ks + vs)
    kv_next = torch.empty_like(kv)
    mk = lambda: torch.empty(kv.shape, dtype=be.acc_dtype, device=kv.device)
    step_g, g_in, g_out = mk(), mk(), None       # this step's contribution / arriving / leaving
    dq_acc = [torch.empty(x.shape, dtype=be.acc_dtype, device=x.device) for x in qs]
    started = [False] * nch
    for s, (src, pairs) in enumerate(ring_schedule(ring.rank, ring.world, causal)):
        last = s + 1 == ring.world
        h_kv = ring.start(kv, kv_next) if not last else []
        h_g = ring.start(g_out, g_in) if s > 0 else []
        touched = [False] * (2 * nch)
        for qc, kc, cz in pairs:
            dq_p, dk_p, dv_p = be.attn_bwd(dOs[qc], o_ch[qc], lse[qc], qs[qc], kv[kc], kv[nch + kc], cz)
            be.accumulate(dq_acc[qc], dq_p, not started[qc])
            started[qc] = True
            be.accumulate(step_g[kc], dk_p, not touched[kc])
            be.accumulate(step_g[nch + kc], dv_p, not touched[nch + kc])
            touched[kc] = touched[nch + kc] = True
        for c, t in enumerate(touched):
            if not t:
                step_g[c].zero_()
        ring.wait(h_kv)
        ring.wait(h_g)
        if s > 0:
            be.accumulate(g_in, step_g, False)       # travelling gradient += this rank's share
            g_out, g_in = g_in, g_out
        else:
            g_out, step_g = step_g, mk()
            if g_in is None:
                g_in = mk()
        if h_kv:
            kv, kv_next = kv_next, kv
    # after the last step the gradient block sits one hop before its owner
    if ring.world > 1:
        ring.wait(ring.start(g_out, g_in))
        dkv = g_in
    else:
        dkv = g_out
    dq = torch.empty_like(dO)
    dk = torch.empty_like(torch.cat(ks, dim=2)) if nch > 1 else torch.empty_like(ks[0])
    dv = torch.empty_like(dk)
    c = qs[0].shape[2]
    for i in range(nch):
        be.store_rows(dq, dq_acc[i], i * c)
        be.store_rows(dk, dkv[i], i * c)
        be.store_rows(dv, dkv[nch + i], i * c)
    return dq, dk, dv


class _RingAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, group, backend):
        o, res = ring_attention_forward(q, k, v, causal=causal, group=group, backend=backend)
        ctx.res, ctx.meta = res, (causal, group, backend)
        return o

    @staticmethod
    def backward(ctx, dO):
        causal, group, backend = ctx.meta
        dq, dk, dv = ring_attention_backward(dO.contiguous(), ctx.res, causal=causal, group=group,
                                             backend=backend)
        return dq, dk, dv, None, None, None


def ring_flash_attention(q, k, v, *, causal: bool, group=None, backend=None):
    """`flash_attention` over a sequence sharded across the ranks of `group` (differentiable).
    Causal inputs must be in zig-zag order (`zigzag_shard`)."""
    return _RingAttentionFn.apply(q, k, v, bool(causal), group, backend)


# ------------------------------------------------------------------------------------------
# single-process form: all ranks' shards in one process, K / V over NVLink P2P inside the library
# (nnop_ring_attn_fwd / nnop_ring_attn_bwd, include/nnop_b200.h) -- what a Julia host calls
# ------------------------------------------------------------------------------------------
import ctypes as _C


def _ptr_array(ts):
    return (_C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _p2p_common(qs, ks, vs):
    W = len(qs)
    if not (len(ks) == len(vs) == W and W >= 1):
        raise NNopError(6, "ring attention: q, k, v must be lists of one tensor per rank")
    CudaBackend._need_cuda(*qs, *ks, *vs)
    B, QH, Ll, E = qs[0].shape
    KH = ks[0].shape[1]
    for q, k, v in zip(qs, ks, vs):
        if tuple(q.shape) != (B, QH, Ll, E) or tuple(k.shape) != (B, KH, Ll, E) or k.shape != v.shape:
            raise NNopError(1, "ring attention: every rank must hold equally shaped shards")
        if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
            raise NNopError(6, "ring attention: shards must be contiguous")
    devs = (_C.c_int * W)(*[q.device.index for q in qs])
    streams = (_C.c_void_p * W)(*[torch.cuda.current_stream(q.device).cuda_stream for q in qs])
    return W, B, QH, KH, Ll, E, devs, streams


def p2p_ring_attention_forward(qs, ks, vs, *, causal: bool):
    """Forward over per-rank shards ``qs[r] (B,QH,Ll,E)``, ``ks[r] / vs[r] (B,KH,Ll,E)`` living on (possibly
    repeated) CUDA devices of this process; causal inputs in zig-zag order.  Returns ``(os, lses)``."""
    W, B, QH, KH, Ll, E, devs, streams = _p2p_common(qs, ks, vs)
    dt = ops._dt(qs[0])
    os_ = [torch.empty_like(q) for q in qs]
    lses = [torch.empty(B, QH, Ll, dtype=torch.float32, device=q.device) for q in qs]
    nbytes = lib.nnop_ring_attn_fwd_workspace_bytes(dt, E, Ll, QH, KH, B, W, int(causal))
    wss = [torch.empty(max(nbytes, 1), dtype=torch.uint8, device=q.device) for q in qs]
    check(lib.nnop_ring_attn_fwd(_ptr_array(os_), _ptr_array(lses), _ptr_array(qs), _ptr_array(ks), _ptr_array(vs),
                                 devs, W, dt, E, Ll, QH, KH, B, int(causal), 1.0 / E ** 0.5, _ptr_array(wss), nbytes,
                                 streams))
    return os_, lses


def p2p_ring_attention_backward(dOs, os_, lses, qs, ks, vs, *, causal: bool):
    """Backward of `p2p_ring_attention_forward`; returns per-rank ``(dqs, dks, dvs)``."""
    W, B, QH, KH, Ll, E, devs, streams = _p2p_common(qs, ks, vs)
    dt = ops._dt(qs[0])
    dqs = [torch.empty_like(q) for q in qs]
    dks = [torch.empty_like(k) for k in ks]
    dvs = [torch.empty_like(v) for v in vs]
    nbytes = lib.nnop_ring_attn_bwd_workspace_bytes(dt, E, Ll, QH, KH, B, W, int(causal))
    wss = [torch.empty(max(nbytes, 1), dtype=torch.uint8, device=q.device) for q in qs]
    dOs = [d.contiguous() for d in dOs]
    check(lib.nnop_ring_attn_bwd(_ptr_array(dqs), _ptr_array(dks), _ptr_array(dvs), _ptr_array(dOs), _ptr_array(os_),
                                 _ptr_array(lses), _ptr_array(qs), _ptr_array(ks), _ptr_array(vs), devs, W, dt, E, Ll,
                                 QH, KH, B, int(causal), 1.0 / E ** 0.5, _ptr_array(wss), nbytes, streams))
    return dqs, dks, dvs
