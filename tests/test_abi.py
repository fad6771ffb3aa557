"""The C-ABI library loads without a GPU, exports every symbol include/scope_ffi.h declares,
and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:scope|b200)_[a-z0-9_]+)\s*\(", text)))


def test_scope_ffi_exports_all_declared_symbols(pkg):
    lib = pkg._ffi.load()
    declared = _declared("scope_ffi.h")
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libscope_b200.so does not export {name}"
    assert sorted(pkg._ffi.EXPORTED_SYMBOLS) == declared
    assert lib.scope_abi_version() == 1
    assert lib.scope_wave_bytes(3840) == 256 * 3840 * 4


def test_cm_shim_exports_all_declared_symbols(pkg):
    lib = C.CDLL(pkg._ffi.SHIM_PATH)
    declared = _declared("cm_shim.h")
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libcm_shim.so does not export {name}"


def test_product_does_not_touch_the_oracle():
    """nothing under the package (or the C sources) references oracle/"""
    pkgdir = os.path.join(ROOT, "obs-color-monitor_b200")
    for base, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_no_gpu_means_loud_failure(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ScopeError) as e:
        pkg.ScopeEngine()
    assert e.value.code == pkg._ffi.SCOPE_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
