"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/kprn_b200.h declares,
the Python binding and the LuaJIT shim declare the same set, and the product fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "kprn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kprn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from kprn_b200 import engine
    lib = ctypes.CDLL(engine.LIB_PATH)
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libkprn_b200.so does not export " + n


def test_python_binding_covers_the_header():
    from kprn_b200 import engine
    declared = set(header_functions()) - {"kprn_last_error", "kprn_version"}
    assert declared == set(engine.ABI.keys())


def test_lua_shim_cdef_matches_header():
    lua = open(os.path.join(ROOT, "kprn_b200", "lua", "kprn_b200.lua")).read()
    for n in header_functions():
        assert n in lua, "LuaJIT shim does not bind " + n


def test_no_cpu_fallback_create_fails_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kprn_b200 import engine
    with pytest.raises(engine.KprnError) as ei:
        engine.Engine(T=6, d_t=4, d_e=4, d_r=4, H=4, L=46, V_e=10)
    assert ei.value.code == 2   # KPRN_ECUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "kprn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".lua", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), "%s mentions the oracle" % f
