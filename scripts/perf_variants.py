"""Time fwd (and optionally bwd) of config C2 for each experiment variant library (development aid).
usage: python scripts/perf_variants.py [--bwd] name1 name2 ...   (names from scripts/build_variants.py)"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, torch
sys.path.insert(0, r"%s")
import nnop_b200 as nn
bwd = %d
torch.manual_seed(0)
def err():
    B,H,L,E = 1,2,1024,128
    q,k,v,dO = (torch.randn(B,H,L,E,device="cuda",dtype=torch.bfloat16) for _ in range(4))
    o,lse = nn._flash_attention(q,k,v,causal=True)
    qf,kf,vf = (t.float().requires_grad_(True) for t in (q,k,v))
    ref = torch.nn.functional.scaled_dot_product_attention(qf,kf,vf,is_causal=True)
    e = [(o.float()-ref).abs().max().item()]
    if bwd:
        g = torch.autograd.grad(ref,(qf,kf,vf),dO.float())
        d = nn.grad_flash_attention(dO,o,lse,q,k,v,causal=True)
        e += [(a.float()-b).abs().max().item() for a,b in zip(d[:3],g)]
    return e
def t(B,H,L,E,causal,iters=8,KH=None):
    KH = KH or H
    q,dO = (torch.randn(B,H,L,E,device="cuda",dtype=torch.bfloat16) for _ in range(2))
    k,v = (torch.randn(B,KH,L,E,device="cuda",dtype=torch.bfloat16) for _ in range(2))
    f = 4.0*B*H*L*L*E*(0.5 if causal else 1)
    for _ in range(3): o,lse = nn._flash_attention(q,k,v,causal=causal)
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters): o,lse = nn._flash_attention(q,k,v,causal=causal)
    b.record(); torch.cuda.synchronize()
    tf = a.elapsed_time(b)/iters
    s = "fwd %%.3f ms %%.0f TF/s" %% (tf, f/tf/1e9)
    if bwd:
        for _ in range(2): nn.grad_flash_attention(dO,o,lse,q,k,v,causal=causal)
        torch.cuda.synchronize(); a.record()
        for _ in range(iters): nn.grad_flash_attention(dO,o,lse,q,k,v,causal=causal)
        b.record(); torch.cuda.synchronize()
        tb = a.elapsed_time(b)/iters
        s += " | bwd %%.3f ms %%.0f TF/s | tot %%.0f TF/s" %% (tb, 2.5*f/tb/1e9, 3.5*f/(tf+tb)/1e9)
    return s
print("err", ["%%.4f" %% x for x in err()], "| C2:", t(8,32,8192,128,True), "| noncausal:", t(4,32,8192,128,False), "| L2048:", t(8,32,2048,128,True), "| GQA32/8:", t(4,32,8192,128,True,KH=8), flush=True)
'''
bwd = "--bwd" in sys.argv
for name in [a for a in sys.argv[1:] if not a.startswith("--")]:
    lib = ROOT / "nnop.jl_b200" / "lib" / ("libnnop_b200.so" if name == "default" else f"variants/libnnop_b200_{name}.so")
    env = dict(os.environ, NNOP_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT / "nnop.jl_b200", int(bwd))], env=env, capture_output=True, text=True, timeout=600)
    print(f"{name:12s}", (r.stdout.strip() or r.stderr.strip()[-600:]), flush=True)
