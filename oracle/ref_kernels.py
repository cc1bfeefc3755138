"""TEST INFRASTRUCTURE (never imported by the product package).

A second, independent restatement of the reference -- this time of its FUSED KERNELS, tile loop by tile loop,
not of the naive functions its tests compare against (those are `oracle/oracle.py`).  NumPy, float64, one
vectorised statement per `@unroll` loop of a workgroup; `gsz` is the reference's workgroup size (= tile size).
File:line citations are into /root/reference (pxl-th/NNop.jl v0.2.0).

Why it exists: the reference ships no golden vectors and cannot run here (no Julia), so the oracle is unpinned
at bit level.  What CAN be pinned on CPU is that three readings of the reference agree with each other:
    its kernels (this file)  ==  its naive test functions (oracle.py)  ==  the residual convention of the
    new library (one lse = m + log l instead of (ms, ls)),
on the reference's own test shapes, including the ragged-tile guards (`in_seq_bounds`), the GQA head mapping
`cld(q_head, n_q_per_kv)`, the top-left causal mask, the key padding mask and `pair` -- see
tests/test_oracle_vs_reference_kernels.py.  Deviation kept out on purpose: the reference's backward stages Q and K
in Float16 shared memory for every T (src/attention_bwd.jl:19-20); `stage_f16=True` reproduces that rounding so
its size can be measured, the default keeps float64.
"""
from __future__ import annotations

import math

import numpy as np

NEG_INF = -np.inf


def _cld(a, b):
    return -(-a // b)


# ------------------------------------------------------------------------------------------------
# flash attention forward: `_flash_attention_fwd!`, src/attention.jl:1-131
# Arrays are in the reference's column-major shapes: q (E, QL, QH, B), k / v (E, KL, KH, B),
# pair (QH, QL, KL, B), kpad_mask (KL, B) bool.  Returns o (E, QL, QH, B), ms, ls (QL, QH, B).
# ------------------------------------------------------------------------------------------------
def flash_attention_fwd(q, k, v, pair=None, kpad_mask=None, *, causal: bool, gsz: int):
    E, QL, QH, B = q.shape
    _, KL, KH, _ = k.shape
    scale = 1.0 / math.sqrt(E)                       # src/attention.jl:154
    n_q_per_kv = QH // KH                            # :14
    kv_seq_tiles = _cld(KL, gsz)                     # :13
    o = np.zeros_like(q, dtype=np.float64)
    ms = np.zeros((QL, QH, B))
    ls = np.zeros((QL, QH, B))
    for b in range(B):
        for qh in range(QH):                         # gidx[2] (0-based here)
            kvh = _cld(qh + 1, n_q_per_kv) - 1       # :28  kv_head_idx = cld(q_head_idx, n_q_per_kv)
            for g in range(_cld(QL, gsz)):           # gidx[1]
                q_off = g * gsz                      # :24
                rows = np.arange(gsz) + q_off        # tidx + q_offset
                in_q = rows < QL                     # :25
                q_shm = np.where(in_q[:, None], q[:, np.minimum(rows, QL - 1), qh, b].T, 0.0)   # :38 (gsz, E)
                o_shm = np.zeros((E, gsz))           # :39-41
                l_i = np.zeros(gsz)                  # :44
                m_i = np.full(gsz, NEG_INF)          # :45
                end_iter = (g + 1) if causal else kv_seq_tiles      # :47
                k_off = 0
                for _ in range(end_iter):
                    cols = np.arange(gsz) + k_off
                    in_k = cols < KL                                                            # :50
                    k_shm = np.where(in_k[None, :], k[:, np.minimum(cols, KL - 1), kvh, b], 0.0)   # :51 (E, gsz)
                    s = (q_shm @ k_shm) * scale                                                 # :55
                    valid = in_q[:, None] & in_k[None, :]
                    if pair is not None:                                                        # :59-64
                        pt = pair[qh][np.ix_(np.minimum(rows, QL - 1), np.minimum(cols, KL - 1))][:, :, b]
                        s = np.where(valid, s + pt, s)
                    if causal:                                                                  # :67-72
                        s = np.where(in_k[None, :] & (rows[:, None] < cols[None, :]), NEG_INF, s)
                    if kpad_mask is not None:                                                   # :73-79
                        keep = np.where(in_k, kpad_mask[np.minimum(cols, KL - 1), b], True)
                        s = np.where(in_k[None, :] & ~keep[None, :], NEG_INF, s)
                    # online softmax over the in-bounds columns only (the loops `break` past KL): :82-94
                    s_in = np.where(in_k[None, :], s, NEG_INF)
                    m_ij = s_in.max(axis=1)
                    with np.errstate(invalid="ignore"):
                        p = np.where(in_k[None, :], np.exp(s - m_ij[:, None]), s)               # out-of-range columns keep s
                        l_ij = np.where(in_k[None, :], p, 0.0).sum(axis=1)
                        m_new = np.maximum(m_i, m_ij)                                           # :97
                        alpha = np.exp(m_i - m_new)                                             # :98
                        beta = np.exp(m_ij - m_new)                                             # :99
                        l_new = alpha * l_i + beta * l_ij                                       # :100
                        p_scale = beta / l_new                                                  # :102
                        o_scale = l_i / l_new * alpha                                           # :103
                        p = p * p_scale[:, None]                                                # :105-107
                        o_shm = o_shm * o_scale[None, :]                                        # :108-110
                        v_shm = np.where(in_k[None, :], v[:, np.minimum(cols, KL - 1), kvh, b], 0.0)   # :113
                        o_shm = o_shm + v_shm @ p.T                                             # :115 (V rows past KL are 0)
                    m_i, l_i = m_new, l_new                                                     # :118-119
                    k_off += gsz
                sel = rows[in_q]
                o[:, sel, qh, b] = o_shm[:, in_q]                                               # :124-130
                ms[sel, qh, b] = m_i[in_q]
                ls[sel, qh, b] = l_i[in_q]
    return o, ms, ls


# ------------------------------------------------------------------------------------------------
# backward: `_flash_attention_bwd_preprocess!` (src/attention_bwd.jl:163-197) + `_flash_attention_bwd!` (:1-161)
# ------------------------------------------------------------------------------------------------
def flash_attention_bwd(d_o, o, ms, ls, q, k, v, pair=None, kpad_mask=None, *, causal: bool, gsz: int,
                        stage_f16: bool = False):
    E, QL, QH, B = q.shape
    _, KL, KH, _ = k.shape
    scale = 1.0 / math.sqrt(E)
    n_q_per_kv = QH // KH
    # preprocess: Δ' = Δ / ls ; δ = Σ_e Δ'·o          (:182-196)
    d_s = d_o / ls[None]
    delta = (d_s * o).sum(axis=0)
    dq = np.zeros_like(q, dtype=np.float64)          # KA.zeros, :233-235
    dk = np.zeros_like(k, dtype=np.float64)
    dv = np.zeros_like(v, dtype=np.float64)
    dpair = np.zeros_like(pair, dtype=np.float64) if pair is not None else None
    f16 = (lambda a: a.astype(np.float16).astype(np.float64)) if stage_f16 else (lambda a: a)
    for b in range(B):
        for qh in range(QH):                         # one workgroup per (head, batch), :262-263
            kvh = _cld(qh + 1, n_q_per_kv) - 1       # :28
            for sn in range(_cld(KL, gsz)):          # key tiles, :39
                lo_k = sn * gsz
                cols = np.arange(gsz) + lo_k
                in_k = cols < KL                                                                # :43
                k_shm = f16(np.where(in_k[None, :], k[:, np.minimum(cols, KL - 1), kvh, b], 0.0))   # :44 (E, gsz)
                for sm in range(sn if causal else 0, _cld(QL, gsz)):                            # :47-48
                    lo_q = sm * gsz
                    rows = np.arange(gsz) + lo_q
                    in_q = rows < QL                                                            # :52
                    rq = np.minimum(rows, QL - 1)
                    d_shm = np.where(in_q[None, :], d_s[:, rq, qh, b], 0.0)                     # :53  (E, gsz)
                    q_shm = f16(np.where(in_q[:, None], q[:, rq, qh, b].T, 0.0))                # :54  (gsz, E)
                    s = (q_shm @ k_shm) * scale                                                 # :58
                    valid = in_q[:, None] & in_k[None, :]
                    if pair is not None:                                                        # :63-68
                        pt = pair[qh][np.ix_(rq, np.minimum(cols, KL - 1))][:, :, b]
                        s = np.where(valid, s + pt, s)
                    if causal:                                                                  # :71-76
                        s = np.where(in_k[None, :] & (rows[:, None] < cols[None, :]), NEG_INF, s)
                    if kpad_mask is not None:                                                   # :77-83
                        keep = np.where(in_k, kpad_mask[np.minimum(cols, KL - 1), b], True)
                        s = np.where(in_k[None, :] & ~keep[None, :], NEG_INF, s)
                    m_i = np.where(in_q, ms[rq, qh, b], np.inf)                                 # :86-87
                    p = np.exp(s - m_i[:, None])                                                # :88-90 (un-normalised P~)
                    dv_t = d_shm @ p                                                            # :94   dV tile (E, gsz keys)
                    dv[:, cols[in_k], kvh, b] += dv_t[:, in_k]                                  # :96-105 (+= / atomic)
                    v_shm = np.where(in_k[None, :], v[:, np.minimum(cols, KL - 1), kvh, b], 0.0)   # :108
                    d_i = np.where(in_q, delta[rq, qh, b], 0.0)                                 # :113-117
                    ds = p * ((d_shm.T @ v_shm) - d_i[:, None]) * scale                         # :111-119
                    if pair is not None:                                                        # :123-132
                        blk = (ds / scale)[np.ix_(in_q, in_k)]
                        dpair[qh, rows[in_q][:, None], cols[in_k][None, :], b] = blk
                    dk_t = q_shm.T @ ds                                                         # :134  (E, gsz keys)
                    dk[:, cols[in_k], kvh, b] += dk_t[:, in_k]                                  # :136-144
                    dq_t = k_shm @ ds.T                                                         # :150  (E, gsz queries)
                    dq[:, rows[in_q], qh, b] += dq_t[:, in_q]                                   # :147-156
    return dq, dk, dv, dpair


# ------------------------------------------------------------------------------------------------
# online softmax: `online_softmax!` + `md_reduce`, src/softmax.jl:1-58 (x (N, cols), softmax over dim 1)
# ------------------------------------------------------------------------------------------------
def _md_reduce(a, b):
    (am, ad), (bm, bd) = a, b
    big, small = ((am, ad), (bm, bd)) if am > bm else ((bm, bd), (am, ad))    # :7-8
    diff = small[0] - big[0]
    if math.isnan(diff):                                                       # :11
        diff = NEG_INF
    return big[0], big[1] + small[1] * math.exp(diff)                          # :12-15


def online_softmax(x, *, gsz: int = 256):
    N, cols = x.shape
    y = np.empty_like(x, dtype=np.float64)
    for c in range(cols):
        partial = [(NEG_INF, 0.0)] * gsz
        for idx in range(gsz):                                                 # each thread's strided walk, :33-40
            md = (NEG_INF, 0.0)
            for e in range(idx, N, gsz):
                md = _md_reduce(md, (float(x[e, c]), 1.0))
            partial[idx] = md
        md = partial[0]
        for other in partial[1:]:                                              # @groupreduce md_reduce, :43
            md = _md_reduce(md, other)
        y[:, c] = np.exp(x[:, c] - md[0]) / md[1]                              # :48-57
    return y


# ------------------------------------------------------------------------------------------------
# RMS norm: `_rms_norm!` src/rms_norm.jl:3-38, `_∇rms_norm!` :43-115 (+ sum over partial rows, :166)
# x (emb, n), w (emb)
# ------------------------------------------------------------------------------------------------
def rms_norm_fwd(x, w, *, eps=1e-6, offset=0.0):
    emb = x.shape[0]
    rstd = 1.0 / np.sqrt((x ** 2).sum(axis=0) * (1.0 / emb) + eps)            # :16-27
    return (offset + w)[:, None] * x * rstd[None, :], rstd                     # :31-36


def rms_norm_bwd(dy, rstd, x, w, *, offset=0.0, batches_per_group=4):
    emb, n = x.shape
    dx = np.empty_like(x, dtype=np.float64)
    groups = _cld(n, batches_per_group)                                        # :141-146
    dw = np.zeros((groups, emb))
    for bid in range(groups):
        for i in range(bid * batches_per_group, min(n, (bid + 1) * batches_per_group)):   # :69-71
            dd = (dy[:, i] * (w + offset) * x[:, i]).sum()                     # :73-83
            m = dy[:, i] * (w + offset)                                        # :95
            dx[:, i] = rstd[i] * m + rstd[i] * (-(1.0 / emb) * rstd[i] ** 2 * dd * x[:, i])   # :96
            dw[bid] += dy[:, i] * (x[:, i] * rstd[i])                          # :98, :101
    return dx, dw.sum(axis=0)                                                  # sum(dw; dims=1), :166


# ------------------------------------------------------------------------------------------------
# layer norm: `_layer_norm!` src/layer_norm.jl:8-63, `_∇layer_norm!` :65-148
# ------------------------------------------------------------------------------------------------
def layer_norm_fwd(x, w, b, *, eps=1e-6):
    emb = x.shape[0]
    mu = x.sum(axis=0) * (1.0 / emb)                                           # :21-30
    var = ((x - mu[None]) ** 2).sum(axis=0) * (1.0 / emb)                      # :36-46 (biased, about the mean)
    rstd = 1.0 / np.sqrt(var + eps)                                            # :48
    return (x - mu[None]) * rstd[None] * w[:, None] + b[:, None], mu, rstd     # :54-61


def layer_norm_bwd(dy, mu, rstd, x, w, *, batches_per_group=4):
    emb, n = x.shape
    dx = np.empty_like(x, dtype=np.float64)
    groups = _cld(n, batches_per_group)
    dw = np.zeros((groups, emb))
    db = np.zeros((groups, emb))
    for bid in range(groups):
        for i in range(bid * batches_per_group, min(n, (bid + 1) * batches_per_group)):   # :92-94
            xn = (x[:, i] - mu[i]) * rstd[i]                                   # :103
            wdy = w * dy[:, i]                                                 # :104
            c1 = (wdy * xn).sum() * (1.0 / emb)                                # :105, :110
            c2 = wdy.sum() * (1.0 / emb)                                       # :106, :111
            dx[:, i] = (wdy - (xn * c1 + c2)) * rstd[i]                        # :128
            dw[bid] += dy[:, i] * xn                                           # :129, :132
            db[bid] += dy[:, i]                                                # :133
    return dx, dw.sum(axis=0), db.sum(axis=0)                                  # sum(...; dims=1), :201


# ------------------------------------------------------------------------------------------------
# Llama RoPE: `llama_rope!`, src/rope/llama_rope.jl:24-65 (in place on copies, :75-76); bwd = sin_sign -1 (:86)
# q (E, L, QH, B), k (E, L, KH, B), cos / sin (E, L, B): only rows 1..E/2 are read (:43-44)
# ------------------------------------------------------------------------------------------------
def llama_rope(q, k, cos, sin, *, sin_sign=1.0):
    q, k = q.astype(np.float64).copy(), k.astype(np.float64).copy()
    half = q.shape[0] // 2
    for x in (q, k):
        for h in range(x.shape[2]):
            for b in range(x.shape[3]):
                c = cos[:half, :, b]
                s = sin[:half, :, b] * sin_sign                                # :44
                x1, x2 = x[:half, :, h, b].copy(), x[half:, :, h, b].copy()    # :47-48
                x[:half, :, h, b] = x1 * c - x2 * s                            # :50
                x[half:, :, h, b] = x2 * c + x1 * s                            # :51
    return q, k
