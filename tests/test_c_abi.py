"""The C-ABI boundary (no compute calls: runs without a GPU): libtsidb.so loads, exports every function that
include/tsidb.h declares, the ctypes struct mirrors have the sizes the header implies, and entry points fail
loudly (negative return + message) instead of falling back when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as ge
from tsid_control_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    return _capi.load_library()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "tsidb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsidb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_entry_point_is_exported(lib):
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/tsidb.h but not exported by libtsidb.so"
    assert sorted(_capi.EXPORTED_SYMBOLS) == names, "tsid_control_b200/_capi.py and include/tsidb.h disagree"


def test_struct_mirrors_match_the_header_layout():
    B, NA = 24, 23
    assert C.sizeof(_capi.TsidbModel) == 8 * ((4 + 4 * B + 7) // 8) + 8 * (9 * B + 3 * B + B + 3 * B + 9 * B) + 8 + 8 * (18 + 6 + 3)
    assert C.sizeof(_capi.TsidbRefs) == 6 * 8 and C.sizeof(_capi.TsidbAuxOut) == 6 * 8
    assert C.sizeof(_capi.TsidbGaitConf) == 5 * 8
    fixed = 12 + 3 + 3 + 6 + 6 + 1 + 6 + 1 + 6 + 6 + 1 + 3 + 3 + 1 + 2 * NA + 1 + 3  # doubles up to kp_am
    assert C.sizeof(_capi.TsidbConf) == 8 * (fixed + 1 + 2 * NA + 1 + 2 * NA + 1 + 1 + 1)


def test_no_cpu_fallback(lib):
    """Without a CUDA device tsidb_create must fail (< 0) with a message; with one this test is skipped."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    from common import setup

    s = setup("v1")
    h = C.c_void_p()
    rc = lib.tsidb_create(C.byref(s["cm"]), C.byref(s["cc"]), 16, 0, C.byref(h))
    assert rc < 0 and not h.value
    assert b"no CUDA device" in lib.tsidb_last_error()
    out = C.c_double()
    assert lib.tsidb_fp64_peak(0, C.byref(out)) < 0
