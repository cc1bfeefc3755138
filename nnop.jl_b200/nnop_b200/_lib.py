"""ctypes binding of libnnop_b200.so (the C ABI declared in include/nnop_b200.h).

This is the Python twin of the `ccall` layer in julia/NNopB200/src/NNopB200.jl: it only
passes device pointers, sizes and the current CUDA stream.  There is no CPU or PyTorch
fallback -- if the shared library is missing, import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("NNOP_B200_LIB", _HERE.parent / "lib" / "libnnop_b200.so"))

NNOP_F32, NNOP_F16, NNOP_BF16 = 0, 1, 2


class NNopError(RuntimeError):
    """Raised when a libnnop_b200 call returns a non-zero status (Julia shim: `error(msg)`)."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class DeviceInfo(C.Structure):
    _fields_ = [("sm_count", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("shared_mem_per_block_optin", C.c_size_t), ("l2_bytes", C.c_size_t),
                ("hbm_bytes", C.c_size_t)]


if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python nnop.jl_b200/build.py` "
        "(nvcc, sm_100a).  nnop_b200 has no CPU fallback.")

lib = C.CDLL(str(LIB_PATH))

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_pp, _ip = C.POINTER(C.c_void_p), C.POINTER(C.c_int)   # host arrays: per-rank device pointers / device ids
# name -> (restype, argtypes); must list every symbol include/nnop_b200.h and nnop_b200_diag.h declare
SIGNATURES = {
    "nnop_version": (_i, []),
    "nnop_last_error_string": (C.c_char_p, []),
    "nnop_device_info": (_i, [_i, C.POINTER(DeviceInfo)]),
    "nnop_set_attention_path": (_i, [_i]),
    "nnop_last_attention_path": (_i, []),
    "nnop_set_bwd_pair_mode": (_i, [_i]),
    "nnop_set_fwd_mode": (_i, [_i]),
    "nnop_flash_attn_fwd": (_i, [_vp] * 7 + [_i] * 8 + [_f, _vp]),
    "nnop_flash_attn_fwd_workspace_bytes": (_sz, [_i] * 7),
    "nnop_flash_attn_pair_workspace_bytes": (_sz, [_i] * 6),
    "nnop_flash_attn_fwd_ws": (_i, [_vp] * 7 + [_i] * 8 + [_f, _vp, _sz, _vp]),
    "nnop_flash_attn_bwd_workspace_bytes": (_sz, [_i] * 7),
    "nnop_flash_attn_bwd": (_i, [_vp] * 12 + [_i] * 8 + [_f, _vp, _sz, _vp]),
    "nnop_flash_attn_bwd_reuse_pair": (_i, [_vp] * 12 + [_i] * 8 + [_f, _vp, _sz, _vp, _vp]),
    "nnop_flash_attn_varlen_fwd": (_i, [_vp] * 7 + [_i] * 3 + [_i64, _i64] + [_i] * 5 + [_f, _vp]),
    "nnop_flash_attn_varlen_fwd_workspace_bytes": (_sz, [_i, _i, _i, _i64, _i]),
    "nnop_flash_attn_varlen_fwd_ws": (_i, [_vp] * 7 + [_i] * 3 + [_i64, _i64] + [_i] * 5 + [_f, _vp, _sz, _vp]),
    "nnop_flash_attn_varlen_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i64, _i]),
    "nnop_flash_attn_varlen_bwd": (_i, [_vp] * 11 + [_i] * 3 + [_i64, _i64] + [_i] * 5 + [_f, _vp, _sz, _vp]),
    "nnop_attn_merge": (_i, [_vp] * 5 + [_i, _i, _i64, _i, _vp]),
    "nnop_accumulate_f32": (_i, [_vp, _vp, _i, _i64, _i, _vp]),
    "nnop_store_rows_from_f32": (_i, [_vp, _vp, _i, _i] + [_i64] * 4 + [_vp]),
    "nnop_ring_attn_fwd_workspace_bytes": (_sz, [_i] * 8),
    "nnop_ring_attn_fwd": (_i, [_pp] * 5 + [_ip] + [_i] * 8 + [_f, _pp, _sz, _pp]),
    "nnop_ring_attn_bwd_workspace_bytes": (_sz, [_i] * 8),
    "nnop_ring_attn_bwd": (_i, [_pp] * 9 + [_ip] + [_i] * 8 + [_f, _pp, _sz, _pp]),
    "nnop_softmax_fwd": (_i, [_vp, _vp, _i, _i64, _i64, _vp]),
    "nnop_softmax_bwd": (_i, [_vp, _vp, _vp, _i, _i64, _i64, _vp]),
    "nnop_rms_norm_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i64, _f, _f, _vp]),
    "nnop_norm_bwd_workspace_bytes": (_sz, [_i64, _i64]),
    "nnop_rms_norm_bwd": (_i, [_vp] * 6 + [_i, _i64, _i64, _f, _vp, _sz, _vp]),
    "nnop_layer_norm_fwd": (_i, [_vp] * 6 + [_i, _i64, _i64, _f, _vp]),
    "nnop_layer_norm_bwd": (_i, [_vp] * 8 + [_i, _i64, _i64, _vp, _sz, _vp]),
    "nnop_llama_rope": (_i, [_vp] * 6 + [_i, _i, _i64, _i, _i, _i, _f, _vp]),
    "nnop_set_timing_events": (_i, [_i, _vp, _vp]),
    "nnop_selftest_umma": (_i, [_vp, _vp, _vp, _i, _vp]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def check(status: int) -> None:
    if status != 0:
        msg = lib.nnop_last_error_string()
        raise NNopError(status, (msg or b"").decode() or f"libnnop_b200 error {status}")
