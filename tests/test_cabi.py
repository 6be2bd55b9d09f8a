"""The C-ABI library: loads, exports every symbol include/locomouse_b200.h declares, and fails loudly
without a GPU (no CPU fallback).  No compute calls here — CPU only."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "locomouse_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*\s*|int64_t\s+|int\s+)(lm_[a-z_0-9]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_entry_points():
    from locomouse_cpp_b200 import api

    assert _declared_functions() == sorted(api.EXPORTS)


def test_library_loads_and_exports_every_declared_symbol():
    from locomouse_cpp_b200 import api

    L = api.load_library()
    for name in _declared_functions():
        assert hasattr(L, name), name
    assert L.lm_abi_version() == 5


def test_struct_layouts_match_the_header():
    """ctypes mirrors must have the C sizes (x86-64 SysV): lm_cand 16, lm_template 24, lm_config 72, lm_results 96."""
    from locomouse_cpp_b200 import types

    assert ctypes.sizeof(types.lm_template) == 24
    assert ctypes.sizeof(types.lm_config) == 72
    assert ctypes.sizeof(types.lm_results) == 96
    assert types.CAND_DTYPE.itemsize == 16
    assert ctypes.sizeof(types.lm_location_prior) == 56 and ctypes.sizeof(types.lm_pairwise_params) == 56
    assert ctypes.sizeof(types.lm_bb_de_params) == 56
    assert ctypes.sizeof(types.lm_bb_base_params) == 48


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from locomouse_cpp_b200 import api, synth

    L = api.load_library()
    ctx = ctypes.c_void_p()
    rc = L.lm_create(ctypes.byref(ctx), 0)
    assert rc == -2 and not ctx.value
    assert b"no CPU fallback" in L.lm_last_error(None)
    spec = synth.SynthSpec()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        api.Detector(spec.config(), synth.make_model(spec), synth.make_background(spec), synth.make_calibration(spec))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under locomouse_cpp_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "locomouse_cpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cpp", ".h", ".hpp", "Makefile")):
                continue
            txt = open(os.path.join(dirpath, fn), errors="ignore").read()
            assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), fn
            assert not re.search(r"#include\s+[\"<].*lm_oracle", txt), fn
            assert "liblm_oracle" not in txt and "lmo_" not in txt, fn
