"""BASELINE config 4 timing: packed batch of 64 sequences, L log-uniform in [128, 16384] (seed 0), H=32, E=128,
bf16 causal, fwd + bwd on one GPU (development aid; CUDA events, inputs resident)."""
import math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch
import nnop_b200 as nn
g = torch.Generator().manual_seed(0)
lens = torch.exp(torch.empty(64).uniform_(math.log(128), math.log(16384), generator=g)).round().int().tolist()
H, E = 32, 128
T = sum(lens)
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
q, k, v, dO = (torch.randn(H, T, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
mx = max(lens)
f = sum(4.0 * H * l * l * E * 0.5 for l in lens)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
o, lse = nn._flash_attention_varlen(q, k, v, cu, cu, mx, mx, causal=True)
tfs = {}
for mode in (1, 2, 1, 2):   # forward variants: one CTA per q tile / persistent, alternating
    nn.set_fwd_mode(mode)
    t_ = timeit(lambda: nn._flash_attention_varlen(q, k, v, cu, cu, mx, mx, causal=True))
    tfs[mode] = min(tfs.get(mode, 1e9), t_)
nn.set_fwd_mode(0)
print("forward: " + "  ".join(f"mode {m}: {t:.3f} ms {f/t/1e9:.0f} TF/s" for m, t in tfs.items()), flush=True)
tf = timeit(lambda: nn._flash_attention_varlen(q, k, v, cu, cu, mx, mx, causal=True))
tbs = {}
for mode in (2, 3, 2, 3):   # backward variants: one CTA per tile / persistent, alternating
    nn.set_bwd_pair_mode(mode)
    t_ = timeit(lambda: nn.grad_flash_attention_varlen(dO, o, lse, q, k, v, cu, cu, mx, mx, causal=True))
    tbs[mode] = min(tbs.get(mode, 1e9), t_)
nn.set_bwd_pair_mode(0)
print("backward: " + "  ".join(f"mode {m}: {t:.3f} ms {2.5*f/t/1e9:.0f} TF/s" for m, t in tbs.items()), flush=True)
tb = timeit(lambda: nn.grad_flash_attention_varlen(dO, o, lse, q, k, v, cu, cu, mx, mx, causal=True))
print(f"C4: 64 seqs, sum L = {T}, sum L^2 = {sum(l*l for l in lens):.3e}, max L = {mx}, min L = {min(lens)}: "
      f"fwd {tf:.3f} ms {f/tf/1e9:.0f} TF/s | bwd {tb:.3f} ms {2.5*f/tb/1e9:.0f} TF/s | fwd+bwd {3.5*f/(tf+tb)/1e9:.0f} TF/s", flush=True)
# the same tokens as a padded dense batch would cost 64 * max L: report the saving
print(f"padded dense equivalent would be {64 * mx * mx / sum(l*l for l in lens):.1f}x the FLOPs", flush=True)
