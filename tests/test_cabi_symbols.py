"""The C-ABI library builds, loads and exports every entry point include/vb_api.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vb_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from vox_serve_b200 import _lib

    return _lib.load(build_if_missing=True)


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("vb_paged_attn", "vb_gemm_bf16", "vb_sample", "vb_rmsnorm", "vb_rope", "vb_kv_append",
                 "vb_snac_pwconv", "vb_pcm16", "vb_last_error"):
        assert must in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(lib):
    from vox_serve_b200 import _lib

    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in vb_api.h but not exported by libvoxb200.so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in vox_serve_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(declared_symbols())


def test_error_reporting_without_gpu(lib):
    assert lib.vb_version() >= 100
    # argument validation happens before any CUDA call
    rc = lib.vb_rmsnorm(None, None, None, 1, 8, ctypes.c_float(1e-5), 0, None)
    assert rc != 0 and b"null" in lib.vb_last_error()
    rc = lib.vb_gemm_bf16(None, None, None, None, 1, 1, 1, 1, 0, 1, 128, 0, 0, None, None)
    assert rc != 0


def test_product_path_has_no_oracle_import():
    pkg = os.path.join(ROOT, "vox_serve_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)
