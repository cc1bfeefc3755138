#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the fp64 oracle (seeded inputs, tiny shapes).

    python tests/golden/make_golden.py

The reference ships no fixtures and cannot run here (no Julia), so these vectors pin the
ORACLE against accidental edits and give the GPU tests fixed inputs; they are not outputs of
the reference itself (oracle/oracle.py header: "parity unpinned").
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
f64 = torch.float64


def attention_cases():
    g = torch.Generator().manual_seed(0)
    cases = {}
    for name, (B, QH, KH, QL, KL, E, causal, use_pair, use_mask) in {
        "plain": (2, 2, 2, 37, 53, 16, False, False, False),
        "causal": (2, 2, 2, 45, 45, 32, True, False, False),
        "gqa_causal": (1, 4, 2, 33, 33, 16, True, False, False),
        "pair_mask": (2, 2, 2, 29, 41, 16, False, True, True),
        "causal_pair_mask": (2, 3, 1, 40, 40, 16, True, True, True),
    }.items():
        q = torch.randn(B, QH, QL, E, generator=g, dtype=f64)
        k = torch.randn(B, KH, KL, E, generator=g, dtype=f64)
        v = torch.randn(B, KH, KL, E, generator=g, dtype=f64)
        dO = torch.randn(B, QH, QL, E, generator=g, dtype=f64)
        pair = torch.randn(B, KL, QL, QH, generator=g, dtype=f64) if use_pair else None
        mask = None
        if use_mask:  # the reference tests' mask: all true except the tail of the last batch
            mask = torch.ones(B, KL, dtype=torch.bool)
            mask[-1, -11:] = False
        # round inputs to fp32 so the stored fixture is exactly what the kernels consume
        q, k, v, dO = (t.float().double() for t in (q, k, v, dO))
        if pair is not None:
            pair = pair.float().double()
        o, lse = O.naive_attention(q, k, v, pair, causal=causal, kpad_mask=mask, return_lse=True)
        dq, dk, dv, dpair = O.naive_attention_bwd(dO, q, k, v, pair, causal=causal, kpad_mask=mask)
        d = dict(q=q, k=k, v=v, dO=dO, o=o, lse=lse, dq=dq, dk=dk, dv=dv, causal=torch.tensor(causal))
        if pair is not None:
            d.update(pair=pair, dpair=dpair)
        if mask is not None:
            d.update(kpad_mask=mask)
        cases[name] = d
    return cases


def rowwise_cases():
    g = torch.Generator().manual_seed(1)
    out = {}
    for emb, n in ((15, 3), (256, 5), (513, 2)):
        x = torch.rand(n, emb, generator=g, dtype=f64).float().double()
        w = torch.rand(emb, generator=g, dtype=f64).float().double()
        b = torch.rand(emb, generator=g, dtype=f64).float().double()
        dy = torch.randn(n, emb, generator=g, dtype=f64).float().double()
        tag = f"{emb}x{n}"
        y = O.naive_softmax(x)
        out[f"softmax_{tag}"] = dict(x=x, dy=dy, y=y, dx=O.naive_softmax_bwd(dy, y))
        for off in (0.0, 1.0):
            yr, rstd = O.naive_rms_norm(x, w, eps=1e-6, offset=off, return_rstd=True)
            dx, dw = O.naive_rms_norm_bwd(dy, x, w, eps=1e-6, offset=off)
            out[f"rms_{tag}_off{int(off)}"] = dict(x=x, w=w, dy=dy, y=yr, rstd=rstd, dx=dx, dw=dw,
                                                   offset=torch.tensor(off))
        yl, mu, rs = O.naive_layer_norm(x, w, b, eps=1e-6, return_stats=True)
        dx, dw, db = O.naive_layer_norm_bwd(dy, x, w, b, eps=1e-6)
        out[f"ln_{tag}"] = dict(x=x, w=w, b=b, dy=dy, y=yl, mean=mu, rstd=rs, dx=dx, dw=dw, db=db)
    return out


def rope_cases():
    g = torch.Generator().manual_seed(2)
    out = {}
    for L, QH, KH, E, B in ((13, 3, 1, 16, 1), (33, 4, 2, 32, 2)):
        pos = torch.arange(L, dtype=torch.float32).view(1, L).repeat(B, 1)
        cos, sin = O.llama_rotary_embedding(E, pos)
        q = torch.randn(B, QH, L, E, generator=g, dtype=f64).float().double()
        k = torch.randn(B, KH, L, E, generator=g, dtype=f64).float().double()
        qo, ko = O.naive_llama_rope(q, k, cos=cos.double(), sin=sin.double())
        qb, kb = O.naive_llama_rope(q, k, cos=cos.double(), sin=sin.double(), bwd=True)
        out[f"rope_L{L}"] = dict(q=q, k=k, cos=cos, sin=sin, q_out=qo, k_out=ko, q_bwd=qb, k_bwd=kb)
    return out


def save(fname, cases):
    flat = {}
    for cname, d in cases.items():
        for key, t in d.items():
            a = t.numpy()
            if a.dtype == np.float64:
                a = a.astype(np.float64)
            flat[f"{cname}/{key}"] = a
    np.savez_compressed(OUT / fname, **flat)
    print(fname, sum(a.nbytes for a in flat.values()) // 1024, "KiB")


if __name__ == "__main__":
    save("attention.npz", attention_cases())
    save("rowwise.npz", rowwise_cases())
    save("rope.npz", rope_cases())
