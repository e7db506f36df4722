"""The drop-in claim at the C++ level: oracle/ref_binding_test.cpp specialises
KernelMatrix<double>::calculate / diagonal (common/kernel_matrix.h:67-107) for a kernel type that carries its
parameters, is compiled against the UNMODIFIED reference headers and sources and linked to libstemk_b200.so.  Here the
reference's own CPU KernelMatrix and the bound one run on the same reference-built MData (Data's constructor,
DAGBuilder, find_root ...) and must agree to 1e-9, with KernelMatrix::print text equal."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conftest import TH, relerr  # noqa: E402
from oracle import refbind as R  # noqa: E402
from stem_kernel_b200 import synth  # noqa: E402

BINDING = os.path.join(ROOT, "oracle", "_ref", "libstemk_ref_binding.so")
needs_binding = pytest.mark.skipif(not os.path.exists(BINDING), reason="oracle/_ref/libstemk_ref_binding.so not built")


@pytest.fixture()
def binding_lib():
    R.use_library("libstemk_ref_binding.so")
    yield R.lib()
    R.use_library("libstemk_ref.so")


@needs_binding
def test_binding_library_exports(binding_lib):
    """CPU: the compiled binding loads (it links the product library) and exports its three entry points next to the
    reference harness's."""
    for sym in ("refbind_gram", "refbind_cross", "refbind_diag", "ref_gram", "ref_mdata_new"):
        assert hasattr(binding_lib, sym), sym


@needs_binding
@pytest.mark.gpu
@pytest.mark.parametrize("kind", [R.SU_STEM, R.SU_STEM_STR, R.SI_STEM_STR, R.STR_SUBST])
def test_bound_kernel_matrix_equals_reference(binding_lib, kind):
    recs = synth.make_config(1, 7, offset=900) + synth.make_config(3, 3, offset=910) + \
        [synth.alignment_like(5, i, n_rows=3) for i in range(2)]
    ref = [R.RefMData.build(r["rows"], r["bp"], TH) for r in recs]
    labels = [r["label"] for r in recs]
    k = R.RefKernel(kind, len_band=10)
    for normalize in (False, True):
        want, _, want_text = k.gram(ref, normalize=normalize, labels=labels, want_text=True)
        got, got_text = k.bound_gram(ref, normalize=normalize, labels=labels)
        assert relerr(got, want) < 1e-9
        assert got_text == want_text
    test, train = ref[:4], ref[4:]          # n_test <= n_train (kernel_matrix.cpp:713,721)
    want, want_self, _ = k.cross(test, train, norm_test=True, normalize=True)
    got, got_self = k.bound_cross(test, train, norm_test=True, normalize=True)
    assert relerr(got, want) < 1e-9 and relerr(got_self, want_self) < 1e-9
    want = k.diag(train)[0]
    assert relerr(k.bound_diag(train), want) < 1e-9
    sv = [0, 3, 5]
    got = k.bound_diag(train, sv_index=sv, init=-1.0)
    want = k.diag(train, sv_index=sv, init=-1.0)[0]
    assert relerr(got, want) < 1e-9
