"""Downstream check of SURVEY 8(c): the vendored LIBSVM (compiled unmodified under oracle/_ref) gives the golden
cross-validation targets on the golden Gram matrix, and they do not move under a 1e-9 relative perturbation --
the tolerance the GPU path is held to."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import refbind as R

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_svm.so")),
                                reason="oracle/_ref not built (needs /root/reference)")


def test_cv_targets_reproduce_and_are_stable_under_tolerance():
    z = np.load(os.path.join(GOLDEN, "golden_svm.npz"))
    K, y = z["gram"], z["y"]
    assert np.array_equal(R.svm_cv(K, y, 1.0, 5, 1), z["cv_target"])
    rng = np.random.default_rng(0)
    Kp = K * (1.0 + 1e-9 * rng.uniform(-1, 1, K.shape))
    Kp = (Kp + Kp.T) / 2
    np.fill_diagonal(Kp, 1.0)
    assert np.array_equal(R.svm_cv(Kp, y, 1.0, 5, 1), z["cv_target"])
