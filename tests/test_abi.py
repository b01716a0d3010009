"""The C-ABI library loads without a GPU and exports every symbol include/tealeaf_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tealeaf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from tealeaf_jl_b200 import lib
    assert declared_symbols() == sorted(lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from tealeaf_jl_b200 import lib
    l = lib.load()
    for name in declared_symbols():
        assert hasattr(l, name), name
    assert l.tl_abi_version() == 1
    assert l.tl_comm_blob_size() > 64


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tealeaf_jl_b200 import lib
    l = lib.load()
    ctx = C.c_void_p()
    assert l.tl_create(C.byref(ctx), 16, 16, 2, 100, 0) == lib.TL_ERR_NO_DEVICE
    assert not ctx.value
    from tealeaf_jl_b200.device import DeviceChunk
    with pytest.raises(lib.TeaLeafError):
        DeviceChunk(16, 16)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tealeaf.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle's", "").lower() or \
                    not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src and "tlo_" not in src, f


def test_every_option_is_documented_in_the_header_and_readable():
    """tl_set_option / tl_get_option names (csrc/tl_api.cu) vs the option list in include/tealeaf_b200.h: every knob is
    documented, and every knob that can be set can be read back (the tests assert code paths through the read-back)."""
    import re
    src = open(os.path.join(ROOT, "tealeaf.jl_b200", "csrc", "tl_api.cu")).read()
    hdr = open(os.path.join(ROOT, "include", "tealeaf_b200.h")).read()
    i, j = src.index('extern "C" int tl_set_option'), src.index('extern "C" int tl_get_option')
    set_names = set(re.findall(r'n == "([a-z0-9_]+)"', src[i:j]))
    get_names = set(re.findall(r'n == "([a-z0-9_]+)"', src[j:src.index('extern "C"', j + 10)]))
    doc = hdr[hdr.index("Tuning / A-B knobs"):hdr.index("int tl_get_option")]
    documented = set(re.findall(r"\b([a-z][a-z0-9_]{3,})\b", doc))
    assert len(set_names) > 25 and len(get_names) > len(set_names)
    assert not set_names - documented, sorted(set_names - documented)
    undocumented = {n for n in get_names - documented if not n.startswith(("prof_", "debug_"))}
    assert not undocumented, sorted(undocumented)
    assert not set_names - get_names, sorted(set_names - get_names)
