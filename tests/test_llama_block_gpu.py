"""BASELINE config 3: the Llama-3-8B attention block as the reference's ops compose it --
rms_norm over the hidden dim, (stand-in) q/k/v projections, llama_rope, GQA causal flash attention
with 32 query / 8 kv heads of E=128 -- forward and backward through the autograd wrappers (the
ChainRules rrules' twin) vs the oracle composed the same way in fp64.  Projections are plain
matmuls (library GEMMs are not part of the hot path); hidden is reduced to keep the oracle fast."""
import pytest
import torch

from helpers import max_abs
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _block(ops, x, w_norm, wq, wk, wv, cos, sin, QH, KH, E, *, oracle):
    B, L, hidden = x.shape
    xn = ops["rms_norm"](x.reshape(B * L, hidden), w_norm).reshape(B, L, hidden)
    heads = lambda t, H: t.reshape(B, L, H, E).permute(0, 2, 1, 3).contiguous()
    q, k, v = heads(xn @ wq, QH), heads(xn @ wk, KH), heads(xn @ wv, KH)
    q, k = ops["rope"](q, k, cos, sin)
    return ops["attn"](q, k, v)


@pytest.mark.parametrize("L", [384, 1000])
def test_llama3_attention_block_fwd_bwd(nnop, L):
    B, hidden, QH, KH, E = 2, 512, 32, 8, 128
    g = torch.Generator().manual_seed(L)
    x = torch.randn(B, L, hidden, generator=g)
    w_norm = 1.0 + 0.1 * torch.randn(hidden, generator=g)
    wq = torch.randn(hidden, QH * E, generator=g) / hidden ** 0.5
    wk = torch.randn(hidden, KH * E, generator=g) / hidden ** 0.5
    wv = torch.randn(hidden, KH * E, generator=g) / hidden ** 0.5
    dO = torch.randn(B, QH, L, E, generator=g)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = nnop.LlamaRotaryEmbedding(E)(pos)

    bf = lambda t: t.to(torch.bfloat16).cuda()
    leaves = [bf(t).requires_grad_(True) for t in (x, w_norm, wq, wk, wv)]
    ours = dict(rms_norm=lambda a, w: nnop.rms_norm(a, w),
                rope=lambda q, k, c, s: nnop.llama_rope(q, k, cos=c, sin=s),
                attn=lambda q, k, v: nnop.flash_attention(q, k, v, causal=True))
    o = _block(ours, *leaves, cos.cuda(), sin.cuda(), QH, KH, E, oracle=False)
    assert nnop.last_attention_path() == 1
    grads = torch.autograd.grad(o, leaves, bf(dO))

    # oracle, fp64, from the SAME bf16-rounded leaves
    rl = [t.detach().double().cpu().requires_grad_(True) for t in leaves]
    ref = dict(rms_norm=lambda a, w: O.naive_rms_norm(a, w),
               rope=lambda q, k, c, s: O.naive_llama_rope(q, k, cos=c, sin=s),
               attn=lambda q, k, v: O.naive_attention(q, k, v, causal=True))
    ro = _block(ref, *rl, cos.double(), sin.double(), QH, KH, E, oracle=True)
    rgrads = torch.autograd.grad(ro, rl, bf(dO).double().cpu())

    assert max_abs(o, ro) < 3e-2, "o"          # chain of four bf16 ops: 1.5x the single-op bound
    for name, a, b in zip(("dx", "dw_norm", "dwq", "dwk", "dwv"), grads, rgrads):
        rel = (a.double().cpu() - b).norm().item() / max(b.norm().item(), 1e-12)
        assert rel < 2e-2, (name, rel)         # reference-style norm-wise check (isapprox, rtol)
