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


_MANT = {torch.bfloat16: 7, torch.float16: 10}


def ulp_T(ref, dtype):
    """Spacing of `dtype` at |ref| (element-wise, fp64): 2^(floor(log2|ref|) - mantissa bits)."""
    a = ref.double().abs().clamp_min(2.0 ** -24)
    return torch.exp2(torch.floor(torch.log2(a)) - _MANT[dtype])


def kernel_err(got, ref):
    """max |got - ref| for Float32 results.  For a 16-bit result T the final rounding of the output to
    T alone moves an element by up to one spacing of T at its magnitude (ulp_T), which says nothing
    about the kernel; what is bounded by BASELINE.json's 2e-2 is the error BEYOND that rounding:
    max(|got - ref| - ulp_T(ref), 0), element-wise ("16-bit bound")."""
    if got.numel() == 0:
        return 0.0
    err = (got.double().cpu() - ref.double().cpu()).abs()
    if got.dtype in _MANT:
        err = (err - ulp_T(ref.cpu(), got.dtype)).clamp_min(0.0)
    return err.max().item()
