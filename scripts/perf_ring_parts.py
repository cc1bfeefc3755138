"""Break ring attention time into its parts (development aid; torchrun, 2+ ranks)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch, torch.distributed as dist
import nnop_b200 as nn
from nnop_b200.ring import _Ring, CudaBackend

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
c, H, E = 8192, 32, 128
def T(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
q, k, v, dO = (torch.randn(1, H, c, E, device="cuda", dtype=torch.bfloat16) for _ in range(4))
kv = torch.randn(4, 1, H, c, E, device="cuda", dtype=torch.bfloat16); kv2 = torch.empty_like(kv)
g32 = torch.randn(4, 1, H, c, E, device="cuda"); g32b = torch.empty_like(g32)
ring = _Ring(None); be = CudaBackend()
res = {}
res["send/recv 268 MB bf16 kv"] = T(lambda: ring.wait(ring.start(kv, kv2)))
res["send/recv 537 MB fp32 dkv"] = T(lambda: ring.wait(ring.start(g32, g32b)))
res["fwd full c x c"] = T(lambda: nn._flash_attention(q, k, v, causal=False))
res["fwd causal c x c"] = T(lambda: nn._flash_attention(q, k, v, causal=True))
o, lse = nn._flash_attention(q, k, v, causal=False)
res["bwd full c x c"] = T(lambda: nn.grad_flash_attention(dO, o, lse, q, k, v, causal=False))
oacc = torch.zeros(1, H, c, E, device="cuda"); l2 = lse.clone()
res["merge"] = T(lambda: be.merge(oacc, l2, o, lse, False))
res["accumulate"] = T(lambda: be.accumulate(oacc, o, False))
res["store_rows"] = T(lambda: be.store_rows(o, oacc, 0))
res["split copies (6 x .contiguous)"] = T(lambda: [x[:, :, :c // 2].contiguous() for x in (q, k, v, q, k, v)])
res["stack kv"] = T(lambda: torch.stack([k, k, v, v]))
if rank == 0:
    for n, t in res.items(): print(f"{n:36s} {t:8.3f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
