"""KernelMatrix::print (common/kernel_matrix.cpp:756-770): LIBSVM precomputed-kernel text, ostream default of six
significant digits.  The Python mirror's formatter must be byte-identical to the reference's printed golden text."""
import io

import numpy as np

from stem_kernel_b200 import api


def test_format_matrix_equals_reference_text(golden):
    z = golden["z"]
    labels = ["%+d" % r["label"] for r in golden["recs"]]
    assert api.format_matrix(z["gram_norm_k3_b10"], labels) == str(z["gram_norm_text_k3_b10"])


def test_special_values_and_print():
    m = np.array([[1.0, np.nan, np.inf], [1e-7, 123456789.0, -0.0]])
    km = api.KernelMatrix()
    km.matrix, km.labels = m, ["+1", "-1"]
    buf = io.StringIO()
    km.print(buf)
    assert buf.getvalue() == "+1 0:1 1:1 2:nan 3:inf \n-1 0:2 1:1e-07 2:1.23457e+08 3:-0 \n"
