"""Quick A/B of the CTA-pair backward against the single-CTA backward (development aid)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, nnop_b200 as nn
torch.manual_seed(0)
def run(B, H, KH, QL, KL, causal, dtype=torch.bfloat16, mask=False):
    q = torch.randn(B, H, QL, 128, device="cuda", dtype=dtype); dO = torch.randn_like(q)
    k = torch.randn(B, KH, KL, 128, device="cuda", dtype=dtype); v = torch.randn_like(k)
    m = None
    if mask:
        m = torch.rand(B, KL, device="cuda") > 0.3; m[:, 0] = True
    o, lse = nn._flash_attention(q, k, v, causal=causal, kpad_mask=m)
    nn.set_bwd_pair_mode(0); ref = nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal, kpad_mask=m)[:3]
    nn.set_bwd_pair_mode(1); got = nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal, kpad_mask=m)[:3]
    torch.cuda.synchronize()
    errs = [(a.float() - b.float()).abs().max().item() for a, b in zip(got, ref)]
    mags = [b.float().abs().max().item() for b in ref]
    print(f"B{B} H{H}/{KH} QL{QL} KL{KL} causal={causal} mask={mask}: max|pair - single| dq {errs[0]:.4f} dk {errs[1]:.4f} dv {errs[2]:.4f} (magnitudes {mags[0]:.1f} {mags[1]:.1f} {mags[2]:.1f})", flush=True)
if len(sys.argv) > 1:
    for ql in (128, 256, 384, 512, 640, 768, 1024):
        run(1, 1, 1, ql, 256, False)
    for kl in (256, 512, 768):
        run(1, 1, 1, 256, kl, False)
    run(1, 2, 2, 384, 384, False)
    run(1, 2, 2, 384, 768, False)
    run(2, 1, 1, 384, 256, False)
    sys.exit(0)
for args in [(1, 1, 1, 256, 256, False), (1, 1, 1, 256, 256, True), (1, 2, 2, 512, 512, True), (2, 4, 2, 1024, 1024, True),
             (1, 2, 2, 384, 384, True), (1, 2, 2, 300, 700, False), (1, 2, 1, 1000, 1000, True), (2, 2, 2, 640, 640, False)]:
    run(*args)
run(2, 2, 2, 512, 512, True, mask=True)
run(1, 2, 2, 512, 512, False, dtype=torch.float16)
