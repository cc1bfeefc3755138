"""The C-ABI library loads and exports every symbol include/nnop_b200.h (and nnop_b200_diag.h) declares (no compute,
no GPU needed), and the product package has no route into oracle/."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "nnop_b200.h").read_text() + (ROOT / "include" / "nnop_b200_diag.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nnop_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib_path = ROOT / "nnop.jl_b200" / "lib" / "libnnop_b200.so"
    assert lib_path.exists(), "build with: python nnop.jl_b200/build.py"
    lib = ctypes.CDLL(str(lib_path))
    syms = _declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in nnop_b200.h but not exported"
    lib.nnop_version.restype = ctypes.c_int
    assert lib.nnop_version() == 100


def test_binding_covers_header(nnop):
    from nnop_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_pure_argument_errors_need_no_gpu(nnop):
    """Validation that happens before any CUDA call (same wording as src/attention.jl:141-144)."""
    from nnop_b200 import _lib
    lib = _lib.lib
    # E not a power of two
    rc = lib.nnop_flash_attn_fwd(None, None, None, None, None, None, None, 0, 48, 8, 8, 2, 2, 1, 0, 1.0, None)
    assert rc == 3 and b"power-of-2" in lib.nnop_last_error_string()
    rc = lib.nnop_flash_attn_fwd(None, None, None, None, None, None, None, 0, 64, 8, 8, 6, 4, 1, 0, 1.0, None)
    assert rc == 1 and b"must be divisible by number of KV heads" in lib.nnop_last_error_string()
    rc = lib.nnop_flash_attn_fwd(None, None, None, None, None, None, None, 7, 64, 8, 8, 4, 4, 1, 0, 1.0, None)
    assert rc == 2
    assert lib.nnop_set_attention_path(5) != 0
    assert lib.nnop_set_attention_path(0) == 0


def test_product_package_never_imports_oracle():
    pkg = ROOT / "nnop.jl_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.jl")):
        txt = f.read_text()
        assert "oracle" not in txt.lower() or f.name == "build.py", f


def test_workspace_queries_need_no_gpu(nnop):
    """Pure host arithmetic of the size queries: the pair extension is one (backward: two) head-major
    copies of the bias with the key axis padded to 32, and the backward workspace covers the padded
    statistics, the fp32 dQ accumulator and the persistent kernel's tile counter."""
    from nnop_b200 import _lib
    lib = _lib.lib
    up = lambda n: (n + 255) & ~255
    for dtype, s in ((0, 4), (1, 2), (2, 2)):
        one = up(3 * 4 * 300 * 320 * s)          # B=3, QH=4, QL=300, KL=300 -> KLp=320
        assert lib.nnop_flash_attn_pair_workspace_bytes(dtype, 300, 300, 4, 3, 0) == one
        assert lib.nnop_flash_attn_pair_workspace_bytes(dtype, 300, 300, 4, 3, 1) == 2 * one
    assert lib.nnop_flash_attn_pair_workspace_bytes(9, 300, 300, 4, 3, 0) == 0
    assert lib.nnop_flash_attn_pair_workspace_bytes(0, 0, 300, 4, 3, 0) == 0
    # bf16, E=128, QL=KL=1000, QH=8, KH=2, B=2: delta + 2 x padded stats + dQ accumulator + counter
    need = lib.nnop_flash_attn_bwd_workspace_bytes(2, 128, 1000, 1000, 8, 2, 2)
    assert need == up(2 * 8 * 1000 * 4) + 2 * up(2 * 8 * 1024 * 4) + up(2 * 8 * 1000 * 128 * 4) + 256
    assert lib.nnop_set_bwd_pair_mode(5) != 0 and lib.nnop_set_bwd_pair_mode(103) == 0
    assert lib.nnop_set_bwd_pair_mode(0) == 0
    # forward: 16-bit problems get a 256-byte workspace (tile counter of the persistent kernel), Float32 E = 64
    # the [hi | lo] fp16 copies of q, k, v plus the 256-byte scale block, Float32 of any other E nothing
    assert lib.nnop_flash_attn_fwd_workspace_bytes(2, 128, 1000, 1000, 8, 2, 2) == 256
    assert lib.nnop_flash_attn_fwd_workspace_bytes(0, 64, 1000, 1000, 8, 2, 2) == (2 * 8 * 1000 + 2 * 2 * 2 * 1000) * 256 + 256
    # Float32 E = 128: [hi 128 | lo 128] fp16 copies for the one-tile split forward; E = 256: SIMT, no workspace
    assert lib.nnop_flash_attn_fwd_workspace_bytes(0, 128, 1000, 1000, 8, 2, 2) == (2 * 8 * 1000 + 2 * 2 * 2 * 1000) * 512 + 256
    assert lib.nnop_flash_attn_fwd_workspace_bytes(0, 256, 1000, 1000, 8, 2, 2) == 0
    assert lib.nnop_flash_attn_varlen_fwd_workspace_bytes(2, 128, 3, 1000, 8) == 256
