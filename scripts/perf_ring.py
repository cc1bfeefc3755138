"""Ring attention throughput (BASELINE config 5 shape family); launch with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/perf_ring.py [L_total] [H] [iters]
Reports whole-job TFLOP/s (causal convention 4*H*L^2*E/2 fwd, x2.5 bwd), max over ranks, CUDA events."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, torch.distributed as dist
import nnop_b200 as nn

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16384 * world
H = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
E, B = 128, 1
Ll = L // world
g = torch.Generator(device="cuda").manual_seed(rank)
q, k, v, dO = (torch.randn(B, H, Ll, E, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(4))
def step(ev=None):
    if ev: ev[0].record()
    o, res = nn.ring_attention_forward(q, k, v, causal=True)
    if ev: ev[1].record()
    out = nn.ring_attention_backward(dO, res, causal=True)
    if ev: ev[2].record()
    return out
for _ in range(3): step()          # same allocation pattern as the timed loop (no cudaMalloc inside it)
torch.cuda.synchronize(); dist.barrier()
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
for e in evs: step(e)
torch.cuda.synchronize()
tf = sum(e[0].elapsed_time(e[1]) for e in evs) / iters
tb = sum(e[1].elapsed_time(e[2]) for e in evs) / iters
t = torch.tensor([tf, tb], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    f = 4.0 * B * H * L * L * E * 0.5
    tf, tb = t.tolist()
    print(f"ring C5-family: world={world} L={L} H={H} E={E} bf16 causal: fwd {tf:.2f} ms {f/tf/1e9:.0f} TF/s | "
          f"bwd {tb:.2f} ms {2.5*f/tb/1e9:.0f} TF/s | fwd+bwd {3.5*f/(tf+tb)/1e9:.0f} TF/s whole job "
          f"({3.5*f/(tf+tb)/1e9/world:.0f} per GPU)", flush=True)
dist.barrier(); dist.destroy_process_group()
