#!/usr/bin/env python
"""Host-buffer (end to end) timing of config C2 through HostAttentionPipeline for several chunk sizes.

    python scripts/perf_e2e.py [kv_heads ...]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import nnop_b200 as nn  # noqa: E402

B, H, L, E = 8, 32, 8192, 128
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pin = lambda: torch.randn(B, H, L, E, generator=g, dtype=torch.float32).to(torch.bfloat16).pin_memory()
q, k, v, dO = pin(), pin(), pin(), pin()
out = {n: torch.empty_like(q).pin_memory() for n in ("o", "dq", "dk", "dv")}
flops = 3.5 * 4 * B * H * L * L * E / 2
for kvh in [int(a) for a in sys.argv[1:]] or [32, 16, 8, 4, 2]:
    for ns in (2, 3):
        pipe = nn.HostAttentionPipeline(q.shape, k.shape, torch.bfloat16, causal=True, chunk=1, kv_heads=kvh,
                                        nslots=ns, device=dev)
        for _ in range(2):
            pipe(q, k, v, dO, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            pipe(q, k, v, dO, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"kv_heads/chunk {kvh:3d} slots {ns}: {ms:7.2f} ms/step  {flops / ms / 1e9:7.1f} TFLOP/s  "
              f"{pipe.h2d_bytes / ms / 1e6:5.1f} GB/s each way", flush=True)
        del pipe
