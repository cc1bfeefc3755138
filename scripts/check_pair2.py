import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
torch.manual_seed(0)
B, H, KL = 1, 1, 256
keep = []
for trial in range(40):
    QL = [512, 640, 384, 1024][trial % 4]
    q = torch.randn(B, H, QL, 128, device="cuda", dtype=torch.bfloat16); dO_full = torch.randn_like(q)
    k = torch.randn(B, H, KL, 128, device="cuda", dtype=torch.bfloat16); v = torch.randn_like(k)
    keep.append((q, k, v, dO_full))   # never reuse memory: every trial touches cold pages
    o, lse = nn._flash_attention(q, k, v, causal=False)
    X = trial % (QL // 128) if trial % 3 else -1
    dO = torch.zeros_like(dO_full)
    if X >= 0: dO[:, :, X * 128:(X + 1) * 128] = dO_full[:, :, X * 128:(X + 1) * 128]
    else: dO = dO_full
    nn.set_bwd_pair_mode(0); ref = nn.grad_flash_attention(dO, o, lse, q, k, v, causal=False)[:3]
    nn.set_bwd_pair_mode(1); got = nn.grad_flash_attention(dO, o, lse, q, k, v, causal=False)[:3]
    torch.cuda.synchronize()
    e = lambda a, b: (a.float() - b.float()).abs()
    dq_e = e(got[0], ref[0])[0, 0].amax(dim=1).view(-1, 128).amax(dim=1).tolist()
    dk_e = e(got[1], ref[1])[0, 0].amax(dim=1).view(-1, 128).amax(dim=1).tolist()
    dv_e = e(got[2], ref[2])[0, 0].amax(dim=1).view(-1, 128).amax(dim=1).tolist()
    bad = max(dq_e + dk_e + dv_e) > 1e-3
    if bad:
        # which columns (head-dim) and rows of dk are wrong?
        dke = e(got[1], ref[1])[0, 0]
        rows = (dke.amax(dim=1) > 1e-3).nonzero().flatten().tolist()
        cols = (dke.amax(dim=0) > 1e-3).nonzero().flatten().tolist()
        dve = e(got[2], ref[2])[0, 0]
        vrows = (dve.amax(dim=1) > 1e-3).nonzero().flatten().tolist()
        vcols = (dve.amax(dim=0) > 1e-3).nonzero().flatten().tolist()
        rng = lambda l: f"{l[0]}..{l[-1]} ({len(l)})" if l else "-"
        print(f"trial {trial} QL {QL} dO block {X}: dq/qblk {['%.2f' % x for x in dq_e]} dk/kvblk {['%.2f' % x for x in dk_e]} dv {['%.2f' % x for x in dv_e]} | dk bad rows {rng(rows)} cols {rng(cols)} | dv bad rows {rng(vrows)} cols {rng(vcols)}", flush=True)
    else:
        print(f"trial {trial} QL {QL} dO block {X}: ok", flush=True)
