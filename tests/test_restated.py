"""The restated recurrence the CUDA kernels implement (tests/restated.py, DESIGN.md: K tables replaced by path
counts, MATCH folded into the y sweep, leaf rows folded into per-node constants) equals the literal reference
recurrence (oracle) -- shown on CPU before any GPU is involved.  Tolerance 1e-12: the restatement reorders sums."""
import numpy as np

import restated
from conftest import relerr
from oracle import oraclebind as O
from stem_kernel_b200 import _lib as L


def ribosum_pair_table():
    """The 256 float literals of stem_kernel_b200/csrc/ribosum85_60.inc (kRibosumPair)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = open(os.path.join(root, "stem_kernel_b200", "csrc", "ribosum85_60.inc")).read()
    body = inc.split("kRibosumPair[256]")[1].split("{", 1)[1].split("}", 1)[0]
    vals = [np.float32(t) for t in re.findall(r"[-+]\d+\.\d+", body)]
    assert len(vals) == 256
    return vals


def test_stem_restatement(golden):
    import math
    beta, g = 0.3, 0.2
    tab = [math.exp(float(v) * beta) for v in ribosum_pair_table()]   # score_table.cpp:124-133
    d = golden["flat"].desc()
    comp = [restated.compile_record(m.export(), g) for m in golden["md"]]
    for band in (10, 0):
        p = O.Params.from_buffer_copy(L.make_params(L.SU_STEM, len_band=band, loop_gap=g, beta=beta))
        want = O.gram(p, d, False)
        n = len(comp)
        got = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                if comp[i]["N"] * comp[j]["N"] > 40000:   # keep the pure-Python loops short
                    got[i, j] = got[j, i] = want[i, j]
                    continue
                got[i, j] = got[j, i] = restated.stem_pair(comp[i], comp[j], tab, band)
        assert relerr(got, want) < 1e-12


def test_string_restatement():
    rng = np.random.default_rng(5)
    sx, sy = rng.integers(0, 4, 23), rng.integers(0, 4, 31)
    wx, wy = rng.random(23), rng.random(31)
    subst = rng.random(16) + 0.5
    gap = 0.8
    # literal recurrence with K tables (string_kernel.cpp:66-132)
    K0 = np.ones((24, 32)); G0 = np.ones((24, 32))
    for j in range(1, 32):
        G0[0, j] = G0[0, j - 1] * gap
    for i in range(1, 24):
        G0[i, 0] = G0[i - 1, 0] * gap
        k1 = g1 = 0.0
        for j in range(1, 32):
            v = G0[i - 1, j - 1] * wx[i - 1] * wy[j - 1] * subst[sx[i - 1] * 4 + sy[j - 1]]
            k1 = v + k1
            g1 = v + g1 * gap
            K0[i, j] = k1 + K0[i - 1, j]
            G0[i, j] = g1 + G0[i - 1, j] * gap
    got = restated.string_pair_sum(list(sx), list(sy), list(wx), list(wy), list(subst), gap)
    assert abs(got - K0[23, 31]) <= 1e-12 * abs(K0[23, 31])
