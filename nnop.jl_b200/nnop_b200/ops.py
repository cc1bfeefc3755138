"""Host launchers mirroring NNop.jl's L2/L3 layers over the C ABI.

Reference functions mirrored (file:line in pxl-th/NNop.jl):
  flash_attention / _flash_attention / ∇flash_attention  src/attention_crc.jl:4-31,
                                                          src/attention.jl:133-177,
                                                          src/attention_bwd.jl:199-275
  online_softmax / ∇online_softmax                        src/softmax.jl:60-86
  rms_norm / _rms_norm / ∇rms_norm                        src/rms_norm.jl:117-185
  layer_norm / _layer_norm / ∇layer_norm                  src/layer_norm.jl:150-220
  LlamaRotaryEmbedding / llama_rope / ∇llama_rope         src/rope/llama_rope.jl:1-98

Shapes are the row-major view of the reference's column-major arrays (same bytes):
q (B,QH,QL,E), k/v (B,KH,KL,E), pair (B,KL,QL,QH), kpad_mask (B,KL) bool, lse (B,QH,QL),
x (n,emb), softmax x (cols,N), cos/sin (B,L,E).  Residual contents differ from the reference
where SURVEY.md Appendix C says so (one fp32 `lse` instead of `ms`,`ls`).
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import NNopError, check, lib

_DT = {torch.float32: _lib.NNOP_F32, torch.float16: _lib.NNOP_F16, torch.bfloat16: _lib.NNOP_BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise NNopError(2, f"unsupported element type {t.dtype}; Float32, Float16, BFloat16 only")


def _req(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise NNopError(6, "nnop_b200 operates on CUDA tensors only (there is no CPU path)")
        if not t.is_contiguous():
            raise NNopError(6, "nnop_b200 requires contiguous (dense column-major) arrays")


def _p(t):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def device_info(device: int | None = None) -> dict:
    """Replaces NNop.shared_memory (src/NNop.jl:27-30, ext/NNopCUDAExt.jl:6-9)."""
    info = _lib.DeviceInfo()
    check(lib.nnop_device_info(torch.cuda.current_device() if device is None else device, info))
    return {f: getattr(info, f) for f, _ in info._fields_}


def set_attention_path(mode: int) -> None:
    check(lib.nnop_set_attention_path(mode))


def last_attention_path() -> int:
    return lib.nnop_last_attention_path()


def set_bwd_pair_mode(mode: int) -> None:
    """Backward kernel variant on the tcgen05 path: 0 automatic (persistent kernel for deep tile queues),
    1 experimental CTA pairs (E = 128), 2 one CTA per tile, 3 persistent, 100+n persistent on n CTAs."""
    check(lib.nnop_set_bwd_pair_mode(int(mode)))


def set_fwd_mode(mode: int) -> None:
    """Forward kernel variant on the tcgen05 path: 0 automatic (persistent kernel for deep tile queues),
    1 one CTA per q tile, 2 persistent, 100+n persistent on n CTAs, 3 the two-softmax-warps-per-row experiment."""
    check(lib.nnop_set_fwd_mode(int(mode)))


def selftest_umma(a: torch.Tensor, b: torch.Tensor, which: int) -> torch.Tensor:
    _req(a, b)
    out = torch.empty(128, 128, dtype=torch.float32, device=a.device)
    check(lib.nnop_selftest_umma(out.data_ptr(), a.data_ptr(), b.data_ptr(), which, _stream()))
    return out


# ------------------------------------------------------------------------------------------
# flash attention
# ------------------------------------------------------------------------------------------
def _attn_dims(q, k, v, pair, kpad_mask):
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise NNopError(1, "q, k, v must be 4-dimensional (E, L, H, B) arrays")
    B, QH, QL, QE = q.shape
    KB, KH, KL, KE = k.shape
    # same checks and wording as src/attention.jl:141-144
    if QE != KE:
        raise NNopError(1, f"Embedding dim of Q `{QE}` must be the same as of K `{KE}`.")
    if tuple(k.shape) != tuple(v.shape):
        raise NNopError(1, f"Shapes of K `{tuple(k.shape)}` and V `{tuple(v.shape)}` must be the same.")
    if KB != B:
        raise NNopError(1, f"Batch size of Q `{B}` must be the same as of K `{KB}`.")
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise NNopError(2, "q, k, v must share one element type")
    if pair is not None:
        if tuple(pair.shape) != (B, KL, QL, QH) or pair.dtype != q.dtype:
            raise NNopError(1, f"pair must be a (QH, QL, KL, B) array of {q.dtype}")
    if kpad_mask is not None:
        if tuple(kpad_mask.shape) != (B, KL) or kpad_mask.dtype != torch.bool:
            raise NNopError(1, "kpad_mask must be a (KL, B) Bool matrix")
    return B, QH, QL, QE, KH, KL


def _up256(n: int) -> int:
    return (n + 255) & ~255


def _flash_attention(q, k, v, pair=None, *, causal: bool, kpad_mask=None, keep_pair_copy: bool = False):
    """`_flash_attention` (src/attention.jl:133-177).  Returns ``(o, lse)``; with ``keep_pair_copy`` a
    third residual (the reference returns a 3-tuple too): the head-major copy of `pair` the library
    made in its workspace (or None), which `grad_flash_attention` can take instead of making its own."""
    _req(q, k, v, pair, kpad_mask)
    B, QH, QL, E, KH, KL = _attn_dims(q, k, v, pair, kpad_mask)
    o = torch.empty_like(q)
    lse = torch.empty((B, QH, QL), dtype=torch.float32, device=q.device)
    scale = 1.0 / math.sqrt(E)
    ws_bytes = lib.nnop_flash_attn_fwd_workspace_bytes(_dt(q), E, QL, KL, QH, KH, B)
    if pair is not None:  # head-major copy of the bias: keeps `pair` on the tensor-core path
        ws_bytes = _up256(ws_bytes) + lib.nnop_flash_attn_pair_workspace_bytes(_dt(q), QL, KL, QH, B, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device) if ws_bytes else None
    check(lib.nnop_flash_attn_fwd_ws(_p(o), _p(lse), _p(q), _p(k), _p(v), _p(pair), _p(kpad_mask),
                                     _dt(q), E, QL, KL, QH, KH, B, int(bool(causal)), scale,
                                     _p(ws), ws_bytes, _stream()))
    if keep_pair_copy:
        hm = None
        if pair is not None and last_attention_path() == 1:  # the copy sits behind the base workspace
            hm = ws[_up256(lib.nnop_flash_attn_fwd_workspace_bytes(_dt(q), E, QL, KL, QH, KH, B)):]
        return o, lse, hm
    return o, lse


def grad_flash_attention(dO, o, lse, q, k, v, pair=None, *, causal: bool, kpad_mask=None,
                         pair_head_major=None):
    """`∇flash_attention` (src/attention_bwd.jl:199-275).  Returns ``(dq, dk, dv, dpair|None)``.
    ``pair_head_major``: the third residual of ``_flash_attention(..., keep_pair_copy=True)``."""
    _req(dO, o, lse, q, k, v, pair, kpad_mask)
    B, QH, QL, E, KH, KL = _attn_dims(q, k, v, pair, kpad_mask)
    if tuple(dO.shape) != tuple(q.shape) or dO.dtype != q.dtype:
        raise NNopError(1, "Δ must have the shape and element type of q")
    dq = torch.empty_like(q)
    dk = torch.empty_like(k)
    dv = torch.empty_like(v)
    dpair = torch.empty_like(pair) if pair is not None else None
    ws_bytes = lib.nnop_flash_attn_bwd_workspace_bytes(_dt(q), E, QL, KL, QH, KH, B)
    if pair is not None:  # (head-major pair +) dpair staging (tensor-core path)
        ws_bytes = _up256(ws_bytes) + lib.nnop_flash_attn_pair_workspace_bytes(
            _dt(q), QL, KL, QH, B, 0 if pair_head_major is not None else 1)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=q.device)
    scale = 1.0 / math.sqrt(E)
    check(lib.nnop_flash_attn_bwd_reuse_pair(
        _p(dq), _p(dk), _p(dv), _p(dpair), _p(dO), _p(o), _p(lse), _p(q), _p(k), _p(v), _p(pair),
        _p(kpad_mask), _dt(q), E, QL, KL, QH, KH, B, int(bool(causal)), scale, _p(ws), ws_bytes, _stream(),
        _p(pair_head_major) if pair is not None else None))
    return dq, dk, dv, dpair


class _FlashAttentionFn(torch.autograd.Function):
    """rrule(_flash_attention) (src/attention_crc.jl:16-31): primal is `o`; residuals o, lse, q, k, v, pair."""

    @staticmethod
    def forward(ctx, q, k, v, pair, causal, kpad_mask):
        o, lse, hm = _flash_attention(q, k, v, pair, causal=causal, kpad_mask=kpad_mask, keep_pair_copy=True)
        ctx.save_for_backward(o, lse, q, k, v, pair, kpad_mask, hm)
        ctx.causal = causal
        return o

    @staticmethod
    def backward(ctx, dO):
        o, lse, q, k, v, pair, kpad_mask, hm = ctx.saved_tensors
        dq, dk, dv, dpair = grad_flash_attention(dO.contiguous(), o, lse, q, k, v, pair,
                                                 causal=ctx.causal, kpad_mask=kpad_mask, pair_head_major=hm)
        return dq, dk, dv, dpair, None, None


def flash_attention(q, k, v, pair=None, *, causal: bool, kpad_mask=None):
    """`NNop.flash_attention(q, k, v, pair=nothing; causal, kpad_mask=nothing)` (src/attention_crc.jl:4-14)."""
    return _FlashAttentionFn.apply(q, k, v, pair, bool(causal), kpad_mask)


# ------------------------------------------------------------------------------------------
# packed variable-length flash attention (additive API)
# ------------------------------------------------------------------------------------------
def _varlen_dims(q, k, v, cu_q, cu_k):
    if q.dim() != 3 or k.dim() != 3 or v.dim() != 3:
        raise NNopError(1, "packed q, k, v must be 3-dimensional (E, L_total, H) arrays")
    QH, TQ, E = q.shape
    KH, TK, KE = k.shape
    if E != KE:
        raise NNopError(1, f"Embedding dim of Q `{E}` must be the same as of K `{KE}`.")
    if tuple(k.shape) != tuple(v.shape):
        raise NNopError(1, f"Shapes of K `{tuple(k.shape)}` and V `{tuple(v.shape)}` must be the same.")
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise NNopError(2, "q, k, v must share one element type")
    for c in (cu_q, cu_k):
        if c.dtype != torch.int32 or c.dim() != 1:
            raise NNopError(1, "cu_seqlens must be Int32 vectors of nseq+1 row offsets")
    if cu_q.numel() != cu_k.numel():
        raise NNopError(1, "cu_seqlens_q and cu_seqlens_k must describe the same number of sequences")
    return QH, TQ, E, KH, TK, cu_q.numel() - 1


def _flash_attention_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q: int,
                            max_seqlen_k: int, *, causal: bool):
    """Packed `_flash_attention`: q (QH, total_q, E), k/v (KH, total_k, E) [Julia (E, total, H)],
    cu_seqlens Int32 device vectors.  Returns ``(o, lse)`` with lse (QH, total_q)."""
    _req(q, k, v, cu_seqlens_q, cu_seqlens_k)
    QH, TQ, E, KH, TK, nseq = _varlen_dims(q, k, v, cu_seqlens_q, cu_seqlens_k)
    o = torch.empty_like(q)
    lse = torch.empty((QH, TQ), dtype=torch.float32, device=q.device)
    ws_bytes = lib.nnop_flash_attn_varlen_fwd_workspace_bytes(_dt(q), E, nseq, TQ, QH)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device) if ws_bytes else None
    check(lib.nnop_flash_attn_varlen_fwd_ws(_p(o), _p(lse), _p(q), _p(k), _p(v), _p(cu_seqlens_q),
                                            _p(cu_seqlens_k), nseq, int(max_seqlen_q), int(max_seqlen_k),
                                            TQ, TK, _dt(q), E, QH, KH, int(bool(causal)),
                                            1.0 / math.sqrt(E), _p(ws), ws_bytes, _stream()))
    return o, lse


def grad_flash_attention_varlen(dO, o, lse, q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q: int,
                                max_seqlen_k: int, *, causal: bool):
    """Packed `∇flash_attention`.  Returns ``(dq, dk, dv)``."""
    _req(dO, o, lse, q, k, v, cu_seqlens_q, cu_seqlens_k)
    QH, TQ, E, KH, TK, nseq = _varlen_dims(q, k, v, cu_seqlens_q, cu_seqlens_k)
    if tuple(dO.shape) != tuple(q.shape) or dO.dtype != q.dtype:
        raise NNopError(1, "Δ must have the shape and element type of q")
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ws_bytes = lib.nnop_flash_attn_varlen_bwd_workspace_bytes(_dt(q), E, nseq, TQ, QH)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=q.device)
    check(lib.nnop_flash_attn_varlen_bwd(_p(dq), _p(dk), _p(dv), _p(dO), _p(o), _p(lse), _p(q), _p(k),
                                         _p(v), _p(cu_seqlens_q), _p(cu_seqlens_k), nseq,
                                         int(max_seqlen_q), int(max_seqlen_k), TQ, TK, _dt(q), E, QH,
                                         KH, int(bool(causal)), 1.0 / math.sqrt(E), _p(ws), ws_bytes,
                                         _stream()))
    return dq, dk, dv


class _FlashAttentionVarlenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_q, cu_k, max_q, max_k, causal):
        o, lse = _flash_attention_varlen(q, k, v, cu_q, cu_k, max_q, max_k, causal=causal)
        ctx.save_for_backward(o, lse, q, k, v, cu_q, cu_k)
        ctx.meta = (max_q, max_k, causal)
        return o

    @staticmethod
    def backward(ctx, dO):
        o, lse, q, k, v, cu_q, cu_k = ctx.saved_tensors
        max_q, max_k, causal = ctx.meta
        dq, dk, dv = grad_flash_attention_varlen(dO.contiguous(), o, lse, q, k, v, cu_q, cu_k, max_q,
                                                 max_k, causal=causal)
        return dq, dk, dv, None, None, None, None, None


def flash_attention_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q: int, max_seqlen_k: int,
                           *, causal: bool):
    """`flash_attention` over a packed batch of variable-length sequences (differentiable)."""
    return _FlashAttentionVarlenFn.apply(q, k, v, cu_seqlens_q, cu_seqlens_k, int(max_seqlen_q),
                                         int(max_seqlen_k), bool(causal))


# ------------------------------------------------------------------------------------------
# online softmax
# ------------------------------------------------------------------------------------------
def _softmax_fwd(x):
    _req(x)
    if x.dim() != 2:
        raise NNopError(1, "online_softmax expects a matrix (src/softmax.jl:60)")
    y = torch.empty_like(x)
    cols, N = x.shape
    check(lib.nnop_softmax_fwd(_p(y), _p(x), _dt(x), N, cols, _stream()))
    return y


def grad_online_softmax(dy, y):
    """`∇online_softmax(Δ, y)` (src/softmax.jl:70-80), fused into one kernel."""
    _req(dy, y)
    if dy.shape != y.shape or dy.dtype != y.dtype:
        raise NNopError(1, "Δ must have the shape and element type of y")
    dx = torch.empty_like(y)
    cols, N = y.shape
    check(lib.nnop_softmax_bwd(_p(dx), _p(dy), _p(y), _dt(y), N, cols, _stream()))
    return dx


class _SoftmaxBwdFn(torch.autograd.Function):
    """`∇online_softmax(Δ, y)` as a differentiable node, so that second-order AD works as in the reference,
    whose pullback switches to plain differentiable broadcasts when it is itself being differentiated
    (`within_gradient(y)`, src/softmax.jl:70-74).  First order: the fused kernel.  Its own pullback, for an
    upstream cotangent g of dx = y∘(Δ − s), s = Σ y∘Δ:
        dΔ = y∘(g − a),  a = Σ y∘g      -- the same operator applied to g: the fused kernel again
        dy = g∘(Δ − s) − a·Δ            -- element-wise + two row sums (torch ops stand in for the GPUArrays
                                           broadcasts the reference uses on this path)"""

    @staticmethod
    def forward(ctx, dy, y):
        ctx.save_for_backward(dy, y)
        return grad_online_softmax(dy.contiguous(), y)

    @staticmethod
    def backward(ctx, g):
        dy, y = ctx.saved_tensors
        g = g.contiguous()
        d_dy = grad_online_softmax(g, y)
        yf, gf, df = y.float(), g.float(), dy.float()
        s = (yf * df).sum(dim=-1, keepdim=True)
        a = (yf * gf).sum(dim=-1, keepdim=True)
        return d_dy, (gf * (df - s) - a * df).to(y.dtype)


class _SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = _softmax_fwd(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        if torch.is_grad_enabled() and (dy.requires_grad or y.requires_grad):
            return _SoftmaxBwdFn.apply(dy, y)   # being differentiated again (create_graph=True)
        return grad_online_softmax(dy.contiguous(), y)


def online_softmax(x):
    """`NNop.online_softmax(x)`: softmax over dim 1 of the (N, cols) matrix (src/softmax.jl:60-68)."""
    return _SoftmaxFn.apply(x)


# ------------------------------------------------------------------------------------------
# RMS norm
# ------------------------------------------------------------------------------------------
def _norm_ws(emb, n, device):
    nbytes = lib.nnop_norm_bwd_workspace_bytes(emb, n)
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device), nbytes


def _rms_norm(x, w, *, eps: float = 1e-6, offset: float = 0.0):
    """`_rms_norm` (src/rms_norm.jl:117-137).  Returns ``(y, rstd)``."""
    _req(x, w)
    if x.dim() != 2 or w.dim() != 1 or w.shape[0] != x.shape[1]:
        raise NNopError(1, "rms_norm: x must be (emb, n) and w (emb) (src/rms_norm.jl:119)")
    if w.dtype != x.dtype:
        raise NNopError(2, "x and w must share one element type")
    n, emb = x.shape
    y = torch.empty_like(x)
    rstd = torch.empty(n, dtype=torch.float32, device=x.device)
    check(lib.nnop_rms_norm_fwd(_p(y), _p(rstd), _p(x), _p(w), _dt(x), emb, n, eps, offset, _stream()))
    return y, rstd


def grad_rms_norm(dy, rstd, x, w, *, offset: float = 0.0):
    """`∇rms_norm(Δ, rms, x, w; offset)` (src/rms_norm.jl:139-169).  ``dw`` is Float32 (:146)."""
    _req(dy, rstd, x, w)
    n, emb = x.shape
    dx = torch.empty_like(x)
    dw = torch.empty(emb, dtype=torch.float32, device=x.device)
    ws, nbytes = _norm_ws(emb, n, x.device)
    check(lib.nnop_rms_norm_bwd(_p(dx), _p(dw), _p(dy), _p(rstd), _p(x), _p(w), _dt(x), emb, n,
                                offset, _p(ws), nbytes, _stream()))
    return dx, dw


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps, offset):
        y, rstd = _rms_norm(x, w, eps=eps, offset=offset)
        ctx.save_for_backward(rstd, x, w)
        ctx.offset = offset
        return y

    @staticmethod
    def backward(ctx, dy):
        rstd, x, w = ctx.saved_tensors
        dx, dw = grad_rms_norm(dy.contiguous(), rstd, x, w, offset=ctx.offset)
        return dx, dw.to(w.dtype), None, None


def rms_norm(x, w, *, eps: float = 1e-6, offset: float = 0.0):
    """`NNop.rms_norm(x, w; ϵ=1f-6, offset=0f0)` (src/rms_norm.jl:171-176)."""
    return _RMSNormFn.apply(x, w, float(eps), float(offset))


# ------------------------------------------------------------------------------------------
# layer norm
# ------------------------------------------------------------------------------------------
def _layer_norm(x, w, b, *, eps: float = 1e-6):
    """`_layer_norm` (src/layer_norm.jl:150-170).  Returns ``(y, mean, rstd)``."""
    _req(x, w, b)
    if x.dim() != 2 or w.dim() != 1 or b.dim() != 1 or w.shape[0] != x.shape[1] or b.shape != w.shape:
        raise NNopError(1, "layer_norm: x must be (emb, n) and w, b (emb)")
    if w.dtype != x.dtype or b.dtype != x.dtype:
        raise NNopError(2, "x, w and b must share one element type")
    n, emb = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(n, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n, dtype=torch.float32, device=x.device)
    check(lib.nnop_layer_norm_fwd(_p(y), _p(mean), _p(rstd), _p(x), _p(w), _p(b), _dt(x), emb, n,
                                  eps, _stream()))
    return y, mean, rstd


def grad_layer_norm(dy, mean, rstd, x, w, b=None):
    """`∇layer_norm(Δ, μ, Σ, x, w, b)` (src/layer_norm.jl:172-204).  Returns ``(dx, dw, db)``."""
    _req(dy, mean, rstd, x, w)
    n, emb = x.shape
    dx = torch.empty_like(x)
    dw = torch.empty_like(w)
    db = torch.empty_like(w)
    ws, nbytes = _norm_ws(emb, n, x.device)
    check(lib.nnop_layer_norm_bwd(_p(dx), _p(dw), _p(db), _p(dy), _p(mean), _p(rstd), _p(x), _p(w),
                                  _dt(x), emb, n, _p(ws), nbytes, _stream()))
    return dx, dw, db


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = _layer_norm(x, w, b, eps=eps)
        ctx.save_for_backward(mean, rstd, x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        mean, rstd, x, w = ctx.saved_tensors
        dx, dw, db = grad_layer_norm(dy.contiguous(), mean, rstd, x, w)
        return dx, dw, db, None


def layer_norm(x, w, b, *, eps: float = 1e-6):
    """`NNop.layer_norm(x, w, b; ϵ=1f-6)` (src/layer_norm.jl:206-211)."""
    return _LayerNormFn.apply(x, w, b, float(eps))


# ------------------------------------------------------------------------------------------
# Llama RoPE
# ------------------------------------------------------------------------------------------
class LlamaRotaryEmbedding:
    """`LlamaRotaryEmbedding(dim; base=10000)` and its functor (src/rope/llama_rope.jl:1-22).

    Host-side table construction, Float32 like the reference: ``inv_freq = 1 / base^((0:2:dim-1)/dim)``;
    calling it on ``position_ids`` (B, L) float32 returns ``cos, sin`` of shape (B, L, dim) with
    the half-frequencies duplicated (``vcat(freqs, freqs)``, :20)."""

    def __init__(self, dim: int, *, base: int = 10000):
        self.dim = int(dim)
        self.base = int(base)
        ids = torch.arange(0, dim, 2, dtype=torch.float32) / float(dim)
        self.inv_freq = 1.0 / (torch.tensor(float(base), dtype=torch.float32) ** ids)

    def __call__(self, position_ids: torch.Tensor):
        pos = position_ids.to(torch.float32)
        freqs = pos.unsqueeze(-1) * self.inv_freq.to(pos.device)
        freqs = torch.cat([freqs, freqs], dim=-1)
        return torch.cos(freqs), torch.sin(freqs)


def _llama_rope(q, k, cos, sin, *, bwd: bool):
    """`_llama_rope(q, k, cos, sin; bwd)` (src/rope/llama_rope.jl:69-89), out-of-place."""
    _req(q, k, cos, sin)
    if q.dim() != 4 or k.dim() != 4:
        raise NNopError(1, "q, k must be (E, L, H, B) arrays")
    # the reference's three @asserts (:70-72)
    if q.shape[3] != k.shape[3] or q.shape[2] != k.shape[2] or q.shape[0] != k.shape[0]:
        raise NNopError(1, "llama_rope: q and k must agree in head dim, sequence length and batch")
    if q.dtype != k.dtype:
        raise NNopError(2, "q and k must share one element type")
    B, QH, L, E = q.shape
    KH = k.shape[1]
    if tuple(cos.shape) != (B, L, E) or tuple(sin.shape) != (B, L, E) or \
            cos.dtype != torch.float32 or sin.dtype != torch.float32:
        raise NNopError(1, "cos, sin must be Float32 (dim, seq, batch) arrays")
    qo = torch.empty_like(q)
    ko = torch.empty_like(k)
    check(lib.nnop_llama_rope(_p(qo), _p(ko), _p(q), _p(k), _p(cos), _p(sin), _dt(q), E, L, QH, KH,
                              B, -1.0 if bwd else 1.0, _stream()))
    return qo, ko


def grad_llama_rope(dq, dk, *, cos, sin):
    """`∇llama_rope(dq, dk; cos, sin)` (src/rope/llama_rope.jl:92)."""
    return _llama_rope(dq, dk, cos, sin, bwd=True)


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, cos, sin):
        ctx.save_for_backward(cos, sin)
        return _llama_rope(q, k, cos, sin, bwd=False)

    @staticmethod
    def backward(ctx, dq, dk):
        cos, sin = ctx.saved_tensors
        gq, gk = _llama_rope(dq.contiguous(), dk.contiguous(), cos, sin, bwd=True)
        return gq, gk, None, None


def llama_rope(q, k, *, cos, sin):
    """`NNop.llama_rope(q, k; cos, sin)` -> ``(q′, k′)`` (src/rope/llama_rope.jl:91)."""
    return _RopeFn.apply(q, k, cos, sin)


# ------------------------------------------------------------------------------------------
# host-buffer entry point (additive; bench.py's end-to-end leg)
# ------------------------------------------------------------------------------------------
def set_timing_events(which: int, start: "torch.cuda.Event | None", stop: "torch.cuda.Event | None"):
    """Arm the one-shot measurement hook of include/nnop_b200.h (nnop_set_timing_events)."""
    check(lib.nnop_set_timing_events(which, None if start is None else start.cuda_event,
                                     None if stop is None else stop.cuda_event))


class HostAttentionPipeline:
    """flash_attention forward + backward on HOST (pinned) arrays.

    The (kv-head group, batch) axis is cut into chunks; chunk c+1 is copied host->device while
    chunk c computes and chunk c-1's results (o, dq, dk, dv) are copied device->host, on three CUDA
    streams with `nslots` device staging slots.  Every (kv-head group, batch) unit is independent
    (src/attention.jl:152) and a GQA group is never split, so chunking changes no result.  A chunk
    is `chunk` batch elements, or -- with `kv_heads` -- `kv_heads` kv heads (and their q heads) of one
    batch element: both are contiguous slabs of the (B, H, L, E) arrays.  Small chunks keep the
    pipeline's fill and drain (one chunk in, one chunk out, not overlapped) short against the
    PCIe-bound steady state.  Device staging is allocated once and reused across calls.
    """

    def __init__(self, q_shape, kv_shape, dtype, *, causal: bool, chunk: int = 1, kv_heads: int = 0,
                 nslots: int = 3, device=None):
        self.causal = bool(causal)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.B, self.QH, self.KH = q_shape[0], q_shape[1], kv_shape[1]
        if self.QH % self.KH:
            raise NNopError(1, "number of q heads must be a multiple of the number of kv heads")
        self.g = self.QH // self.KH
        self.kv_heads = self.KH if kv_heads <= 0 else min(kv_heads, self.KH)
        self.chunk = max(1, min(chunk, self.B)) if self.kv_heads == self.KH else 1
        cq = (self.chunk, self.kv_heads * self.g) + tuple(q_shape[2:])
        ck = (self.chunk, self.kv_heads) + tuple(kv_shape[2:])
        mk = lambda shp: torch.empty(shp, dtype=dtype, device=self.dev)
        self.slots = [dict(q=mk(cq), k=mk(ck), v=mk(ck), dO=mk(cq)) for _ in range(max(2, nslots))]
        self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.h2d_bytes = self.d2h_bytes = 0

    def _chunks(self):
        """(batch slice, kv-head slice) of every chunk, batch-major."""
        for b0 in range(0, self.B, self.chunk):
            for h0 in range(0, self.KH, self.kv_heads):
                yield slice(b0, min(self.B, b0 + self.chunk)), slice(h0, min(self.KH, h0 + self.kv_heads))

    def __call__(self, q, k, v, dO, out):
        """q, dO (B,QH,QL,E), k, v (B,KH,KL,E): pinned CPU tensors.  out: dict of pinned CPU tensors
        'o','dq' like q and 'dk','dv' like k, filled in place.  Blocks until the results are on the host."""
        for t in (q, k, v, dO, *out.values()):
            if t.is_cuda or not t.is_pinned():
                raise NNopError(6, "HostAttentionPipeline expects pinned host tensors")
        ns, g = len(self.slots), self.g
        comp_done = [None] * ns
        last = None
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_h2d, self.s_comp, self.s_d2h):
            s.wait_stream(cur)
        self.h2d_bytes = self.d2h_bytes = 0
        for ci, (bs, hs) in enumerate(self._chunks()):
            n, nh = bs.stop - bs.start, hs.stop - hs.start
            qs = slice(hs.start * g, hs.stop * g)
            slot = self.slots[ci % ns]
            with torch.cuda.stream(self.s_h2d):
                if comp_done[ci % ns] is not None:
                    self.s_h2d.wait_event(comp_done[ci % ns])  # slot inputs consumed
                for name, src, sl in (("q", q, qs), ("k", k, hs), ("v", v, hs), ("dO", dO, qs)):
                    part = src[bs, sl]
                    slot[name][:n, :sl.stop - sl.start].copy_(part, non_blocking=True)
                    self.h2d_bytes += part.numel() * part.element_size()
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_h2d)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(ev_in)
                qd, dOd = slot["q"][:n, :nh * g], slot["dO"][:n, :nh * g]
                kd, vd = slot["k"][:n, :nh], slot["v"][:n, :nh]
                o, lse = _flash_attention(qd, kd, vd, causal=self.causal)
                dq, dk, dv, _ = grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=self.causal)
                ev_c = torch.cuda.Event()
                ev_c.record(self.s_comp)
                comp_done[ci % ns] = ev_c
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev_c)
                for name, src, sl in (("o", o, qs), ("dq", dq, qs), ("dk", dk, hs), ("dv", dv, hs)):
                    out[name][bs, sl].copy_(src, non_blocking=True)
                    src.record_stream(self.s_d2h)
                    self.d2h_bytes += src.numel() * src.element_size()
                lse.record_stream(self.s_d2h)
                last = torch.cuda.Event()
                last.record(self.s_d2h)
        cur.wait_stream(self.s_d2h)
        cur.wait_stream(self.s_comp)
        if last is not None:
            last.synchronize()
        return out
