"""CPU oracle for the NNop hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in torch-CPU, the *naive* definitions that the reference's own
test-suite uses to check its fused kernels.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it; the product
package (``nnop.jl_b200/nnop_b200``) never does and fails loudly without its CUDA library.

PARITY STATUS: "parity unpinned" at bit level.  The reference (pxl-th/NNop.jl v0.2.0) is
Julia + KernelAbstractions; there is no ``julia`` binary in this image, every reference
kernel is declared ``cpu=false`` and the reference ships NO golden vectors / fixtures
(its tests draw unseeded ``randn`` and compare against the naive functions restated here).
What pins this oracle instead:
  * the naive definitions, followed line by line (file:line cited on every function);
  * analytic known-answer tests taken from the reference tests' own inputs
    (RoPE on all-ones q/k, ``test/rope_tests.jl:21-56``; see tests/test_oracle.py);
  * an independent implementation (``torch.nn.functional.scaled_dot_product_attention``,
    ``torch.nn.functional.layer_norm`` / ``rms_norm`` / ``softmax``) and torch autograd
    on the fp64 graph for every closed-form gradient;
  * ``oracle/ref_kernels.py``: a NumPy restatement of the reference's FUSED KERNELS, tile loop by
    tile loop, which must -- and does, to 1e-11 (tests/test_oracle_vs_reference_kernels.py) -- equal
    these naive functions on the reference's test shapes: the equality the reference's own suite
    asserts on a GPU.

Layout.  Julia arrays are column-major; a Julia ``(E, L, H, B)`` array is byte-identical
to a contiguous row-major tensor of shape ``(B, H, L, E)``.  Everything here uses the
row-major view:

    q, o, dq, dO  (B, QH, QL, E)        k, v, dk, dv  (B, KH, KL, E)
    lse           (B, QH, QL)           pair, dpair   (B, KL, QL, QH)   (head fastest)
    kpad_mask     (B, KL) bool, True = attend
    x, y (n, emb) ; w, b (emb) ; rstd, mean (n)       softmax x (cols, N), over last dim
    cos, sin      (B, L, E) float32
"""
from __future__ import annotations

import math

import torch

__all__ = [
    "naive_attention", "naive_attention_bwd", "naive_softmax", "naive_softmax_bwd",
    "naive_rms_norm", "naive_rms_norm_bwd", "naive_layer_norm", "naive_layer_norm_bwd",
    "llama_rotary_embedding", "naive_llama_rope", "rotate_half",
]


# ----------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------
def _scores(q, k, pair, causal, kpad_mask):
    """Masked, scaled logits ``(B, QH, QL, KL)``.

    Follows test/attention_testsetup.jl:21-43: GQA head expansion (:23-30, q-head j uses
    kv-head j // (QH/KH), matching the kernel's ``cld(q_head, n_q_per_kv)`` at
    src/attention.jl:28), ``(kᵀ ⊠ q) .* inv(sqrt(E))`` (:32-33), causal mask keeping
    ``k_idx <= q_idx`` (:34-37, same as src/attention.jl:70), key-padding mask added as
    ``log(mask)`` i.e. -Inf where False (:16-19,38-40) and ``pair`` added un-scaled
    (:41-43, src/attention.jl:62).
    """
    B, QH, QL, E = q.shape
    _, KH, KL, _ = k.shape
    assert QH % KH == 0, "Number of query heads must be divisible by number of KV heads"
    g = QH // KH
    if g > 1:
        k = k.repeat_interleave(g, dim=1)
    s = torch.einsum("bhqe,bhke->bhqk", q, k) * (1.0 / math.sqrt(E))
    neg_inf = torch.tensor(float("-inf"), dtype=s.dtype)
    if causal:
        qi = torch.arange(QL).view(QL, 1)
        ki = torch.arange(KL).view(1, KL)
        s = torch.where(ki <= qi, s, neg_inf)
    if kpad_mask is not None:
        s = torch.where(kpad_mask.view(B, 1, 1, KL), s, neg_inf)
    if pair is not None:
        # pair is (B, KL, QL, QH) row-major == Julia (QH, QL, KL, B)
        s = s + pair.permute(0, 3, 2, 1)
    return s


def naive_attention(q, k, v, pair=None, *, causal: bool, kpad_mask=None, return_lse=False,
                    zero_masked_rows=False):
    """``v ⊠ softmax(scores)`` -- test/attention_testsetup.jl:21-45.

    ``zero_masked_rows``: the reference produces NaN for a query row whose keys are all
    masked (``exp(-Inf - -Inf)``, src/attention.jl:91); the B200 kernels define that row as
    0 with ``lse = -inf``.  Off by default = reference behaviour.
    """
    B, QH, QL, E = q.shape
    KH = k.shape[1]
    g = QH // KH
    s = _scores(q, k, pair, causal, kpad_mask)
    m = s.amax(dim=-1, keepdim=True)
    if zero_masked_rows:
        m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m)
    l = e.sum(dim=-1, keepdim=True)
    if zero_masked_rows:
        p = torch.where(l > 0, e / torch.where(l > 0, l, torch.ones_like(l)), torch.zeros_like(e))
    else:
        p = e / l
    vv = v.repeat_interleave(g, dim=1) if g > 1 else v
    o = torch.einsum("bhqk,bhke->bhqe", p, vv)
    if return_lse:
        lse = (m + torch.log(l)).squeeze(-1)
        return o, lse
    return o


def naive_attention_bwd(dO, q, k, v, pair=None, *, causal: bool, kpad_mask=None,
                        zero_masked_rows=False, o=None):
    """Closed-form gradients of :func:`naive_attention` (the maths
    the reference realises at src/attention_bwd.jl:86-156 and :182-196):

        D = rowsum(dO ∘ O);  dV = Pᵀ dO;  dP = dO Vᵀ;  dS = P ∘ (dP − D);
        dpair = dS;  dQ = E^-1/2 dS K;  dK = E^-1/2 dSᵀ Q   (dK, dV summed over a GQA group)

    ``o``: the forward output the pullback was handed.  The reference's backward is a function of
    ``(Δ, o, ms, ls, q, k, v)`` (src/attention_bwd.jl:199-206) and its preprocess kernel takes D from
    THAT ``o`` (src/attention_bwd.jl:182-196), which for 16-bit element types is the rounded forward
    result; pass it to evaluate the same function of the same arguments.  ``None``: D from the exact P V.

    Returns ``(dq, dk, dv, dpair_or_None)``.
    """
    B, QH, QL, E = q.shape
    _, KH, KL, _ = k.shape
    g = QH // KH
    scale = 1.0 / math.sqrt(E)
    s = _scores(q, k, pair, causal, kpad_mask)
    m = s.amax(dim=-1, keepdim=True)
    if zero_masked_rows:
        m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m)
    l = e.sum(dim=-1, keepdim=True)
    if zero_masked_rows:
        p = torch.where(l > 0, e / torch.where(l > 0, l, torch.ones_like(l)), torch.zeros_like(e))
    else:
        p = e / l
    kk = k.repeat_interleave(g, dim=1) if g > 1 else k
    vv = v.repeat_interleave(g, dim=1) if g > 1 else v
    if o is None:
        o = torch.einsum("bhqk,bhke->bhqe", p, vv)
    D = (dO * o.to(dO.dtype)).sum(dim=-1, keepdim=True)
    dv_full = torch.einsum("bhqk,bhqe->bhke", p, dO)
    dP = torch.einsum("bhqe,bhke->bhqk", dO, vv)
    dS = p * (dP - D)
    dq = torch.einsum("bhqk,bhke->bhqe", dS, kk) * scale
    dk_full = torch.einsum("bhqk,bhqe->bhke", dS, q) * scale
    if g > 1:
        dk = dk_full.view(B, KH, g, KL, E).sum(dim=2)
        dv = dv_full.view(B, KH, g, KL, E).sum(dim=2)
    else:
        dk, dv = dk_full, dv_full
    dpair = dS.permute(0, 3, 2, 1).contiguous() if pair is not None else None
    return dq, dk, dv, dpair


# ----------------------------------------------------------------------------------------
# softmax
# ----------------------------------------------------------------------------------------
def naive_softmax(x):
    """Softmax over the last (row-major) dim == Julia ``dims=1``; test/softmax_tests.jl:6-10."""
    mx = x.amax(dim=-1, keepdim=True)
    tmp = torch.exp(x - mx)
    return tmp / tmp.sum(dim=-1, keepdim=True)


def naive_softmax_bwd(dy, y):
    """``dx = y∘Δ − y·Σ(y∘Δ)`` -- src/softmax.jl:70-80."""
    tmp = dy * y
    return tmp - y * tmp.sum(dim=-1, keepdim=True)


# ----------------------------------------------------------------------------------------
# RMS norm
# ----------------------------------------------------------------------------------------
def naive_rms_norm(x, w, *, eps=1e-6, offset=0.0, return_rstd=False):
    """``(w + offset) · x / sqrt(mean(x²) + ϵ)`` -- test/rmsnorm_tests.jl:7-9
    (kernel: src/rms_norm.jl:16-36, residual ``rms`` holds rstd, :27)."""
    rstd = torch.rsqrt((x * x).mean(dim=-1, keepdim=True) + eps)
    y = (w + offset) * x * rstd
    if return_rstd:
        return y, rstd.squeeze(-1)
    return y


def naive_rms_norm_bwd(dy, x, w, *, eps=1e-6, offset=0.0):
    """``dx = r·δ(w+off) − r³·x·Σ(δ(w+off)x)/N``; ``dw = Σ_rows δ·x·r`` -- src/rms_norm.jl:40-42,72-101."""
    N = x.shape[-1]
    rstd = torch.rsqrt((x * x).mean(dim=-1, keepdim=True) + eps)
    wd = dy * (w + offset)
    dd = (wd * x).sum(dim=-1, keepdim=True)
    dx = rstd * wd - rstd ** 3 * x * dd / N
    dw = (dy * x * rstd).sum(dim=0)
    return dx, dw


# ----------------------------------------------------------------------------------------
# layer norm
# ----------------------------------------------------------------------------------------
def naive_layer_norm(x, w, b, *, eps=1e-6, return_stats=False):
    """``(x − μ)/sqrt(σ² + ϵ)·w + b`` with the biased variance -- test/layernorm_tests.jl:7-11
    (kernel: src/layer_norm.jl:21-61; residuals μ and rstd, :50)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    y = (x - mu) * rstd * w + b
    if return_stats:
        return y, mu.squeeze(-1), rstd.squeeze(-1)
    return y


def naive_layer_norm_bwd(dy, x, w, b, *, eps=1e-6):
    """``dx = (wδ − (x̂·mean(wδx̂) + mean(wδ)))·r``; ``dw = Σ δx̂``; ``db = Σ δ`` -- src/layer_norm.jl:95-136."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    xh = (x - mu) * rstd
    wd = dy * w
    c1 = (wd * xh).mean(dim=-1, keepdim=True)
    c2 = wd.mean(dim=-1, keepdim=True)
    dx = (wd - (xh * c1 + c2)) * rstd
    dw = (dy * xh).sum(dim=0)
    db = dy.sum(dim=0)
    return dx, dw, db


# ----------------------------------------------------------------------------------------
# Llama RoPE
# ----------------------------------------------------------------------------------------
def llama_rotary_embedding(dim: int, position_ids, *, base: int = 10000):
    """``LlamaRotaryEmbedding(dim; base)(position_ids)`` -- src/rope/llama_rope.jl:7-22.

    ``inv_freq[j] = base^-(2j/dim)`` evaluated in Float32 exactly as the reference does
    (``ids = (0:2:dim-1)/dim`` in Float32, ``inv.(base .^ ids)``, :8-9), frequencies
    duplicated by ``vcat`` (:20); ``position_ids`` is ``(B, L)`` float32.
    Returns ``cos, sin`` of shape ``(B, L, dim)`` float32.
    """
    position_ids = torch.as_tensor(position_ids, dtype=torch.float32)
    ids = torch.arange(0, dim, 2, dtype=torch.float32) / float(dim)
    inv_freq = 1.0 / (torch.tensor(float(base), dtype=torch.float32) ** ids)
    freqs = position_ids.unsqueeze(-1) * inv_freq            # (B, L, dim/2)
    freqs = torch.cat([freqs, freqs], dim=-1)                # (B, L, dim)
    return torch.cos(freqs), torch.sin(freqs)


def rotate_half(x):
    """``vcat(-x2, x1)`` along E -- test/rope_tests.jl:6-11."""
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def naive_llama_rope(q, k, *, cos, sin, bwd: bool = False):
    """``x·cos + rotate_half(x)·sin`` on q ``(B,QH,L,E)`` and k ``(B,KH,L,E)`` --
    test/rope_tests.jl:13-19.  ``bwd=True`` flips the sign of sin, which is the reference's
    pullback (src/rope/llama_rope.jl:86,92)."""
    c = cos.unsqueeze(1).to(q.dtype)
    s = sin.unsqueeze(1).to(q.dtype)
    if bwd:
        s = -s
    return q * c + rotate_half(q) * s, k * c + rotate_half(k) * s
