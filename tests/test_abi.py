"""The C-ABI library loads and exports every symbol include/stemk.h declares; without a CUDA device its compute
entry points fail loudly (no CPU fallback).  No compute is attempted here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from stem_kernel_b200 import _lib as L
from stem_kernel_b200 import api


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "stemk.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(stemk_[a-z0-9_]+)\s*\(", h)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(L.CUDA_SO)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.EXPORTS) == names      # the Python binding covers the whole header


def test_struct_layout_matches_header():
    assert C.sizeof(L.Params) == 8 + 8 * 8                 # int32 + uint32 + 8 doubles
    from stem_kernel_b200.hostlib import SeqSetDesc
    assert C.sizeof(SeqSetDesc) == 8 + 20 * 8              # uint32 (+pad) + 20 pointers


def test_version_and_argument_errors():
    lib = L.lib()
    assert b"sm_100a" in lib.stemk_version()
    h = C.c_void_p()
    assert lib.stemk_create(C.byref(h), None, 0) == L.ERR_ARG
    bad = L.make_params(99)
    assert lib.stemk_create(C.byref(h), C.byref(bad), 0) == L.ERR_ARG
    assert b"kind" in lib.stemk_last_error(None)


def test_no_device_means_error_not_fallback():
    if L.lib().stemk_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.StemkError, match="no CUDA device"):
        api.Context(L.make_params(L.SU_STEM))
    with pytest.raises(api.StemkError):
        api.SuStemKernel()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "stem_kernel_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "stemk_oracle" not in src and "libstemk_ref" not in src, f
