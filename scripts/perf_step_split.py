"""Where one C2 step goes (development aid): forward, backward total and the backward main kernel alone
(the C ABI's one-shot timing events), so prep + post = backward total - main kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "nnop.jl_b200"))
import torch
import nnop_b200 as nn

B, H, L, E = 8, 32, 8192, 128
q, k, v, dO = (torch.randn(B, H, L, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
o, lse = nn._flash_attention(q, k, v, causal=True)
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(3): nn.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
n = 10
f0, f1, b0, b1 = ev(), ev(), ev(), ev()
ks = [(ev(), ev()) for _ in range(n)]
for a, b in ks: a.record(); b.record()   # materialise the handles
torch.cuda.synchronize()
f0.record()
for _ in range(n): nn._flash_attention(q, k, v, causal=True)
f1.record(); b0.record()
for i in range(n):
    nn.set_timing_events(1, ks[i][0], ks[i][1])
    nn.grad_flash_attention(dO, o, lse, q, k, v, causal=True)
b1.record(); torch.cuda.synchronize()
main = sum(a.elapsed_time(b) for a, b in ks) / n
tot = b0.elapsed_time(b1) / n
print(f"C2 bf16 causal: forward {f0.elapsed_time(f1)/n:.3f} ms | backward total {tot:.3f} ms, main kernel {main:.3f} ms, "
      f"prep + post + gaps {tot-main:.3f} ms ({100*(tot-main)/tot:.1f} %)")

# prep / post alone (measured r02: 0.572 ms against 0.574 ms of traffic at the copy bandwidth -- one 16-byte vector per
# thread is already at the roofline; 4 vectors per thread was slower, 0.72 ms): graph-free event timing of the backward with the main kernel's time subtracted is noisy,
# so also time a tiny-KL problem of the same Q size (main kernel ~ nothing, prep + post unchanged)
ks_, vs_ = k[:, :, :128].contiguous(), v[:, :, :128].contiguous()
o2, lse2 = nn._flash_attention(q, ks_, vs_, causal=False)
for _ in range(3): nn.grad_flash_attention(dO, o2, lse2, q, ks_, vs_, causal=False)
a, b = ev(), ev(); a.record()
for _ in range(n): nn.grad_flash_attention(dO, o2, lse2, q, ks_, vs_, causal=False)
b.record(); torch.cuda.synchronize()
print(f"same Q, KL = 128 (prep + post + a one-column main kernel): {a.elapsed_time(b)/n:.3f} ms; ideal prep + post traffic "
      f"{(q.numel()*2*2 + q.numel()*4 + q.numel()*4 + q.numel()*2)/1e9:.2f} GB = {(q.numel()*14)/6548e9*1e3:.3f} ms at 6 548 GB/s")
