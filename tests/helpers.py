"""Shared helpers for the parity tests."""
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(fname):
    """-> {case: {key: torch tensor}}"""
    z = np.load(GOLDEN / fname)
    out = {}
    for full in z.files:
        case, key = full.split("/", 1)
        out.setdefault(case, {})[key] = torch.from_numpy(z[full])
    return out


def max_abs(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item() if a.numel() else 0.0


def reference_isapprox(a, b, atol, rtol):
    """Julia `isapprox(a, b; atol, rtol)` for arrays: norm-wise."""
    a = a.double().cpu()
    b = b.double().cpu()
    return (a - b).norm().item() <= max(atol, rtol * max(a.norm().item(), b.norm().item()))
