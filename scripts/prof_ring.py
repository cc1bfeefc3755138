"""torch.profiler table of one ring fwd+bwd (development aid; torchrun)."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, torch.distributed as dist
import nnop_b200 as nn
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
L, H, E = 16384 * world, 32, 128
q, k, v, dO = (torch.randn(1, H, L // world, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
def step():
    o, res = nn.ring_attention_forward(q, k, v, causal=True)
    return nn.ring_attention_backward(dO, res, causal=True)
for _ in range(2): step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter(); o, res = nn.ring_attention_forward(q, k, v, causal=True); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
if rank == 0: print(f"fwd host enqueue {1e3*(t1-t0):.2f} ms, until done {1e3*(t2-t0):.2f} ms", flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60), flush=True)
dist.barrier(); dist.destroy_process_group()
