"""Randomised A/B of the persistent backward against the one-CTA-per-tile kernel (dense and packed):
dK / dV must agree bit for bit, dQ up to the order of its fp32 reduce-adds.  Development aid; the
fixed cases live in tests/."""
import random
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import nnop_b200 as nn  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
rng = random.Random(seed)
t_end = time.time() + budget
n = 0
close = lambda a, b: (a.float() - b.float()).abs().max().item() <= 2 ** -7 * max(1.0, b.float().abs().max().item())
try:
    while time.time() < t_end:
        E = rng.choice([64, 128])
        dt = rng.choice([torch.bfloat16, torch.float16])
        causal = rng.random() < 0.6
        KH = rng.choice([1, 2, 3])
        QH = KH * rng.choice([1, 2, 4])
        mode = rng.choice([3, 101, 102, 103, 105, 111, 3])
        if rng.random() < 0.5:   # dense
            B = rng.choice([1, 2, 3])
            QL = rng.choice([1, 64, 127, 128, 129, 300, 512, 700, 1024, 1500])
            KL = QL if causal else rng.choice([QL, 64, 200, 513, 900])
            q, dO = (torch.randn(B, QH, QL, E, device="cuda").to(dt) for _ in range(2))
            k, v = (torch.randn(B, KH, KL, E, device="cuda").to(dt) for _ in range(2))
            o, lse = nn._flash_attention(q, k, v, causal=causal)
            run = lambda: nn.grad_flash_attention(dO, o, lse, q, k, v, causal=causal)
            desc = f"dense B{B} H{QH}/{KH} L{QL}/{KL} E{E} causal={causal} {dt} mode {mode}"
        else:                    # packed
            ns = rng.choice([1, 2, 5, 9])
            lq = [rng.choice([0, 1, 60, 128, 129, 255, 400, 777, 1100]) for _ in range(ns)]
            lk = lq if causal else [rng.choice([l, 0, 100, 300, 600]) for l in lq]
            if sum(lq) == 0 or sum(lk) == 0:
                continue
            cu = lambda ls: torch.tensor([0] + list(torch.tensor(ls).cumsum(0)), dtype=torch.int32, device="cuda")
            cq, ck = cu(lq), cu(lk)
            q, dO = (torch.randn(QH, sum(lq), E, device="cuda").to(dt) for _ in range(2))
            k, v = (torch.randn(KH, sum(lk), E, device="cuda").to(dt) for _ in range(2))
            mq, mk = max(lq), max(lk)
            o, lse = nn._flash_attention_varlen(q, k, v, cq, ck, mq, mk, causal=causal)
            run = lambda: nn.grad_flash_attention_varlen(dO, o, lse, q, k, v, cq, ck, mq, mk, causal=causal) + (None,)
            desc = f"packed H{QH}/{KH} lq={lq} lk={lk} E{E} causal={causal} {dt} mode {mode}"
        nn.set_bwd_pair_mode(2)
        ref = run()
        nn.set_bwd_pair_mode(mode)
        for rep in range(2):
            got = run()
            assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]), "dK/dV differ: " + desc
            assert close(got[0], ref[0]), "dQ differs: " + desc
            assert torch.isfinite(got[0].float()).all(), "non-finite dQ: " + desc
        n += 1
finally:
    nn.set_bwd_pair_mode(0)
print(f"stress_persist seed {seed}: {n} random cases OK")
