import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch, nnop_b200 as nn
from oracle import oracle as O
from test_attention_gpu import _inputs
for QH in (6, 8):
  for L in (255, 256, 257, 512):
    q, k, v, dO, _, _ = _inputs(2, QH, 1, L, L, 64, torch.float32, L + QH)
    qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
    rq, rk, rv, _ = O.naive_attention_bwd(dO.double(), q.double(), k.double(), v.double(), causal=True)
    res = []
    for mode in (0, 1):
        nn.set_attention_path(mode)
        o, lse = nn._flash_attention(qd, kd, vd, causal=True)
        dq, dk, dv, _ = nn.grad_flash_attention(dOd, o, lse, qd, kd, vd, causal=True)
        e = lambda a, b: (a.double().cpu() - b).abs().max().item()
        res.append(f"{'TC  ' if mode == 0 else 'SIMT'} dq {e(dq, rq):.2e} dk {e(dk, rk):.2e} dv {e(dv, rv):.2e}")
    nn.set_attention_path(0)
    # where is the worst dv error?
    print(f"QH{QH} L{L}: " + " | ".join(res) + f" | mags {rq.abs().max():.1f} {rk.abs().max():.1f} {rv.abs().max():.1f}", flush=True)
