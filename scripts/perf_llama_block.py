"""BASELINE config 3 timing: Llama-3-8B attention block pieces on one GPU (development aid):
rms_norm over hidden 4096 on (L*B) rows, llama_rope on q (32 heads) / k (8 heads), GQA causal
flash attention fwd+bwd, E=128, L=8192, bf16.  CUDA events, inputs resident, B in {1, 4}."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch
import nnop_b200 as nn
PK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}

def T(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for B in (1, 4):
    QH, KH, L, E, hidden = 32, 8, 8192, 128, 4096
    dt = torch.bfloat16
    x = torch.randn(B * L, hidden, device="cuda", dtype=dt); w = torch.rand(hidden, device="cuda", dtype=dt)
    q = torch.randn(B, QH, L, E, device="cuda", dtype=dt); k = torch.randn(B, KH, L, E, device="cuda", dtype=dt)
    v = torch.randn_like(k); dO = torch.randn_like(q); dy = torch.randn_like(x)
    pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
    cos, sin = nn.LlamaRotaryEmbedding(E)(pos); cos, sin = cos.cuda(), sin.cuda()
    y, rstd = nn._rms_norm(x, w)
    t_rf = T(lambda: nn._rms_norm(x, w)); t_rb = T(lambda: nn.grad_rms_norm(dy, rstd, x, w))
    t_rope = T(lambda: nn.llama_rope(q, k, cos=cos, sin=sin))
    o, lse = nn._flash_attention(q, k, v, causal=True)
    t_af = T(lambda: nn._flash_attention(q, k, v, causal=True)); t_ab = T(lambda: nn.grad_flash_attention(dO, o, lse, q, k, v, causal=True))
    f = 4.0 * B * QH * L * L * E * 0.5
    b_rf = 2 * x.numel() * 2 + hidden * 2 + 4 * B * L; b_rb = 3 * x.numel() * 2 + hidden * 2 + 4 * B * L + 4 * hidden
    b_rope = 2 * (q.numel() + k.numel()) * 2 + 2 * (E // 2) * L * B * 4
    print(f"C3 B={B}: rms_norm fwd {t_rf*1e3:.0f} us ({b_rf/t_rf/1e6:.0f} GB/s, {100*b_rf/t_rf/1e6/PK['hbm_gbs']:.0f}%) bwd {t_rb*1e3:.0f} us ({b_rb/t_rb/1e6:.0f} GB/s, {100*b_rb/t_rb/1e6/PK['hbm_gbs']:.0f}%) | "
          f"rope fwd {t_rope*1e3:.0f} us ({b_rope/t_rope/1e6:.0f} GB/s, {100*b_rope/t_rope/1e6/PK['hbm_gbs']:.0f}%) | "
          f"attn fwd {t_af:.3f} ms {f/t_af/1e9:.0f} TF/s bwd {t_ab:.3f} ms {2.5*f/t_ab/1e9:.0f} TF/s fwd+bwd {3.5*f/(t_af+t_ab)/1e9:.0f} TF/s "
          f"({100*3.5*f/(t_af+t_ab)/1e9/PK['bf16_tflops']:.0f}% of measured burst bf16)", flush=True)
    # what fusing RoPE into the attention block could save at most = the RoPE launches themselves
    # (forward: one launch on q, k; backward: one more on dq, dk), against the block with them
    t_rope_b = T(lambda: nn.grad_llama_rope(q, k, cos=cos, sin=sin))
    fwd_blk, trn_blk = t_rf + t_rope + t_af, t_rf + t_rope + t_af + t_ab + t_rope_b + t_rb
    print(f"   f2 ceiling B={B}: forward-only block (rms_norm + rope + attention) {fwd_blk*1e3:.0f} us, rope {t_rope*1e3:.0f} us = "
          f"{100*t_rope/fwd_blk:.1f}% | forward+backward block {trn_blk*1e3:.0f} us, rope fwd+bwd {(t_rope+t_rope_b)*1e3:.0f} us = "
          f"{100*(t_rope+t_rope_b)/trn_blk:.1f}% (the backward needs the rotated q, k in HBM: not fusable there)", flush=True)
