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


def test_native_writer_is_byte_identical(golden):
    """stemk_format_rows / stemk_format_values (the streaming writer of SURVEY 8(f) rank 2) against the reference's
    printed golden text and against the Python mirror on awkward values (ties at six digits, denormals, -0, nan, inf)."""
    z = golden["z"]
    labels = ["%+d" % r["label"] for r in golden["recs"]]
    assert api.format_rows(z["gram_norm_k3_b10"], labels) == str(z["gram_norm_text_k3_b10"]).encode()
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.standard_normal(4000) * 10.0 ** rng.integers(-320, 300, 4000),
                           [0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf, 1e-5, 9.9999995e-5, 999999.5, 1e6, 123456.5,
                            0.1, 1 / 3, 2.5e-310, 1.0, 100000, 1234567, 0.000123456789, 5e-324]])
    m = np.resize(vals, (41, 99))
    lab = [str(i - 20) for i in range(41)]
    for threads in (1, 3, 0):
        assert api.format_rows(m, lab, first_cnt=7, n_threads=threads) == \
            "".join(f"{lab[i]} 0:{i + 7} " + "".join(f"{j + 1}:{api._g6(m[i, j])} " for j in range(99)) + "\n"
                    for i in range(41)).encode()
    assert api.format_rows(np.zeros((0, 5)), []) == b""
    assert api.format_values(vals[:50]) == "".join(api._g6(v) + "\n" for v in vals[:50]).encode()
