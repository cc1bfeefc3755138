"""The C-ABI library loads and exports every symbol include/nnop_b200.h declares (no compute,
no GPU needed), and the product package has no route into oracle/."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "nnop_b200.h").read_text()
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
