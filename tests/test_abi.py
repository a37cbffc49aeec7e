"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute calls (no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from instantrestore_b200.build import build_library
    path = build_library()
    return ctypes.CDLL(str(path))


def _declared_symbols():
    text = (ROOT / "include" / "instantrestore_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ir_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("ir_conv_gemm", "ir_shared_attn_fwd", "ir_groupnorm", "ir_layernorm", "ir_adain_coeffs", "ir_concat_freeu",
              "ir_upsample_nearest2x", "ir_latent_in", "ir_latent_out", "ir_last_error_string", "ir_check_device"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/instantrestore_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from instantrestore_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()


def test_version_and_error_string(lib):
    lib.ir_last_error_string.restype = ctypes.c_char_p
    assert lib.ir_version() >= 100
    assert isinstance(lib.ir_last_error_string(), bytes)


def test_null_arguments_are_rejected_without_a_gpu(lib):
    """Argument validation happens before any CUDA call: NULL params return IR_ERR_ARG (-5), never crash."""
    lib.ir_last_error_string.restype = ctypes.c_char_p
    for fn in ("ir_conv_gemm", "ir_shared_attn_fwd", "ir_groupnorm", "ir_layernorm", "ir_adain_coeffs", "ir_concat_freeu"):
        f = getattr(lib, fn)
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        assert f(None, None) == -5, fn
        assert b"NULL" in lib.ir_last_error_string()


def test_struct_layouts_match_header():
    """ctypes Structure field order/names mirror the header's typedef structs."""
    from instantrestore_b200 import _lib
    text = (ROOT / "include" / "instantrestore_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    pairs = {"ir_conv_gemm_params": _lib.ConvGemmParams, "ir_shared_attn_params": _lib.SharedAttnParams,
             "ir_groupnorm_params": _lib.GroupNormParams, "ir_layernorm_params": _lib.LayerNormParams,
             "ir_adain_coeffs_params": _lib.AdainCoeffsParams, "ir_concat_freeu_params": _lib.ConcatFreeuParams}
    for cname, struct in pairs.items():
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname + ";", text).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            names.append(re.findall(r"[A-Za-z_0-9]+", first)[-1])
            names += [re.findall(r"[A-Za-z_0-9]+", r)[-1] for r in rest]
        assert names == [f[0] for f in struct._fields_], cname


def test_product_path_has_no_oracle_import():
    """The shipped package must never import the oracle (or fall back to it)."""
    for p in (ROOT / "instantrestore_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
