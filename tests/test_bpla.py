"""BPLA / local-alignment kernels (bpla_kernel/bpla_kernel.cpp:16-175).

CPU: the plain-C restatement (oracle/stemk_oracle.c) against the golden values of the compiled reference -- bit
for bit -- and, where oracle/_ref exists, against the reference itself on fresh inputs.  GPU: the CUDA kernel
through the C ABI (stemk_bpla_pairs) against the oracle to 1e-9 relative (observed ~1e-14: CUDA's exp and the
re-associated Y scan), on the golden records, on C1/C2-like sets and on the edge cases the DP has (length 1,
lengths around the 32-column chunk, sequences of one base, rectangular x/y sets)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, need_gpu, relerr
from oracle import oraclebind as O
from oracle import refbind as R
from stem_kernel_b200 import bpla, synth

TOL = 1e-9
VARIANTS = [(0, 0), (0, 1), (1, 0), (1, 1)]


def golden_set():
    z = np.load(os.path.join(GOLDEN, "golden_bpla.npz"))
    rows = json.loads(str(z["rows_json"]))
    off = z["col_off"]
    recs = [dict(rows=r, p_left=z["p_left"][off[k]:off[k + 1]], p_right=z["p_right"][off[k]:off[k + 1]],
                 p_unpair=z["p_unpair"][off[k]:off[k + 1]]) for k, r in enumerate(rows)]
    return bpla.BplaSet(recs), z


CUSTOM = dict(gap=-3.0, ext=-0.25, alpha=2.0, beta=0.3, score=np.arange(16.0).reshape(4, 4) / 4 - 1)


def test_restatement_is_bit_identical_to_the_reference_golden():
    s, z = golden_set()
    assert np.array_equal(s.col_off, z["col_off"])
    for no_bp, sw in VARIANTS:
        got = O.bpla_pairs(bpla.make_params(no_bp=no_bp, sw=sw), s, s, z["xi"], z["yi"])
        assert np.array_equal(got, z[f"k_nobp{no_bp}_sw{sw}"]), (no_bp, sw)
    assert np.array_equal(O.bpla_pairs(bpla.make_params(**CUSTOM), s, s, z["xi"], z["yi"]), z["k_custom"])
    # a length-1 record against itself: M = exp(beta*score), result 1 + M (bpla_kernel.cpp:113-114)
    k = [i for i, r in enumerate(s.rows) if r == ["a"]][0]
    p = bpla.make_params(no_bp=True)
    assert O.bpla_pairs(p, s, s, [k], [k])[0] == 1 + np.exp(p.beta * p.score[0])


@pytest.mark.skipif(not os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_bpla.so")), reason="oracle/_ref not built")
def test_restatement_matches_the_compiled_reference_on_fresh_inputs():
    recs = [dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(1, 7, offset=4000) + synth.make_config(3, 2, offset=10)]
    s = bpla.BplaSet(recs)
    xi, yi = np.triu_indices(len(s))
    for no_bp, sw in VARIANTS:
        p = bpla.make_params(no_bp=no_bp, sw=sw)
        assert np.array_equal(O.bpla_pairs(p, s, s, xi, yi), R.bpla_pairs(p, s, s, xi, yi))


def test_gradient_restatement_is_bit_identical_to_the_reference_golden():
    """BPLAKernel::compute_gradients (bpla_kernel.cpp:176-402): value and d/d{alpha, beta, gap, ext} per pair."""
    s, z = golden_set()
    v, g = O.bpla_gradients(bpla.make_params(), s, s, z["xi"], z["yi"])
    assert np.array_equal(v, z["grad_value"]) and np.array_equal(g, z["grad"])
    v, g = O.bpla_gradients(bpla.make_params(**CUSTOM), s, s, z["xi"], z["yi"])
    assert np.array_equal(v, z["grad_value_custom"]) and np.array_equal(g, z["grad_custom"])


def test_gradients_are_the_derivatives_of_the_value():
    """Independent check of the oracle: central differences of compute_gradients' own value in each parameter."""
    s, z = golden_set()
    xi, yi = z["xi"][:20], z["yi"][:20]
    base = dict(gap=-8.0, ext=-0.75, alpha=4.5, beta=0.11)
    _, g = O.bpla_gradients(bpla.make_params(**base), s, s, xi, yi)
    for k, name in enumerate(("alpha", "beta", "gap", "ext")):
        h = 2.0 ** -12
        lo, hi = dict(base), dict(base)
        lo[name] -= h; hi[name] += h
        pl, ph = bpla.make_params(**lo), bpla.make_params(**hi)     # make_params rounds to float like the reference's CLI
        vl, _ = O.bpla_gradients(pl, s, s, xi, yi)
        vh, _ = O.bpla_gradients(ph, s, s, xi, yi)
        fd = (vh - vl) / (getattr(ph, name) - getattr(pl, name))
        assert np.allclose(fd, g[:, k], rtol=1e-3, atol=1e-6 * np.abs(g).max()), name


@pytest.mark.skipif(not os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_bpla.so")), reason="oracle/_ref not built")
def test_gradient_restatement_matches_the_compiled_reference_on_fresh_inputs():
    recs = [dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(1, 6, offset=4100)]
    s = bpla.BplaSet(recs)
    xi, yi = np.triu_indices(len(s))
    p = bpla.make_params(gap=-5.0, ext=-0.5, alpha=3.0, beta=0.2)
    v, g = O.bpla_gradients(p, s, s, xi, yi)
    rv, rg = R.bpla_gradients(p, s, s, xi, yi)
    assert np.array_equal(v, rv) and np.array_equal(g, rg)


def test_pairing_profiles_follow_the_reference_front_end():
    """data.cpp:19-46: p_left / p_right are square roots of the summed pairing probabilities, p_unpair of the rest."""
    bp = (np.array([1, 1, 3]), np.array([8, 9, 7]), np.array([0.5, 0.25, 0.81]))
    pl, pr, pu = bpla.pairing_profiles(10, bp)
    assert pl.dtype == np.float32 and np.isclose(pl[0] ** 2, 0.75) and np.isclose(pl[2] ** 2, 0.81) and pl[7] == 0
    assert np.isclose(pr[7] ** 2, 0.5) and np.isclose(pr[8] ** 2, 0.25) and np.isclose(pr[6] ** 2, 0.81)
    assert np.allclose(pl ** 2 + pr ** 2 + pu ** 2, 1.0, atol=1e-6)


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_golden_and_edge_cases():
    need_gpu()
    from stem_kernel_b200 import api, _lib as L
    ctx = api.Context(L.make_params(L.STR_SIMPLE))
    s, z = golden_set()
    for no_bp, sw in VARIANTS:
        got = bpla.pairs(ctx, bpla.make_params(no_bp=no_bp, sw=sw), s, s, z["xi"], z["yi"])
        assert relerr(got, z[f"k_nobp{no_bp}_sw{sw}"]) < TOL, (no_bp, sw)
    assert relerr(bpla.pairs(ctx, bpla.make_params(**CUSTOM), s, s, z["xi"], z["yi"]), z["k_custom"]) < TOL
    # lengths around the 32-column chunk, both roles; the kernel is not symmetric in rounding only
    rng = np.random.default_rng(5)
    recs = []
    for n in (1, 2, 31, 32, 33, 64, 65, 97):
        a, b = rng.uniform(0, 0.6, n), rng.uniform(0, 0.4, n)
        recs.append(dict(rows=["".join(rng.choice(list("acgu"), n))], p_left=np.sqrt(a), p_right=np.sqrt(b),
                         p_unpair=np.sqrt(np.maximum(0, 1 - a - b))))
    e = bpla.BplaSet(recs)
    xi, yi = np.divmod(np.arange(len(e) ** 2), len(e))
    for no_bp, sw in VARIANTS:
        p = bpla.make_params(no_bp=no_bp, sw=sw)
        assert relerr(bpla.pairs(ctx, p, e, e, xi, yi), O.bpla_pairs(p, e, e, xi, yi)) < TOL
    with pytest.raises(api.StemkError, match="out of range"):
        bpla.pairs(ctx, bpla.make_params(), e, e, [len(e)], [0])


@pytest.mark.gpu
def test_gpu_gram_on_config1_and_rectangular_config3():
    need_gpu()
    from stem_kernel_b200 import api, _lib as L
    ctx = api.Context(L.make_params(L.STR_SIMPLE))
    c1 = bpla.BplaSet([dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(1, 60)])
    p = bpla.make_params()
    g = bpla.gram(ctx, p, c1, normalize=True)
    iu = np.triu_indices(len(c1))
    want = O.bpla_pairs(p, c1, c1, iu[0], iu[1])
    d = want[iu[0] == iu[1]]
    assert relerr(g[iu], np.where(iu[0] == iu[1], 1.0, want / np.sqrt(d[iu[0]] * d[iu[1]]))) < TOL
    assert np.array_equal(g, g.T) and np.all(np.diag(g) == 1.0)
    # 150-300 nt records: several chunks per row, x and y sets differ
    a = bpla.BplaSet([dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(3, 6, offset=50)])
    b = bpla.BplaSet([dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(3, 5, offset=90)])
    xi, yi = np.divmod(np.arange(30), 5)
    for no_bp, sw in VARIANTS:
        q = bpla.make_params(no_bp=no_bp, sw=sw)
        assert relerr(bpla.pairs(ctx, q, a, b, xi, yi), O.bpla_pairs(q, a, b, xi, yi)) < TOL


@pytest.mark.gpu
def test_gpu_gradients_golden_edge_cases_and_long_records():
    """stemk_bpla_gradients against the reference's golden values and the oracle: every value and every derivative
    within 1e-9 relative (the derivative with respect to beta mixes signs, so its error is measured against the sum
    of the other three magnitudes as well -- observed ~1e-14)."""
    need_gpu()
    from stem_kernel_b200 import api, _lib as L
    ctx = api.Context(L.make_params(L.STR_SIMPLE))
    s, z = golden_set()

    def check(got, want):
        (v, g), (wv, wg) = got, want
        assert relerr(v, wv) < TOL
        scale = np.maximum(np.abs(wg), 1e-3 * np.abs(wg).max(axis=1, keepdims=True))
        assert np.max(np.abs(g - wg) / scale) < TOL

    check(bpla.gradients(ctx, bpla.make_params(), s, s, z["xi"], z["yi"]), (z["grad_value"], z["grad"]))
    check(bpla.gradients(ctx, bpla.make_params(**CUSTOM), s, s, z["xi"], z["yi"]), (z["grad_value_custom"], z["grad_custom"]))
    rng = np.random.default_rng(6)
    recs = []
    for n in (1, 2, 31, 32, 33, 64, 65, 97):
        a, b = rng.uniform(0, 0.6, n), rng.uniform(0, 0.4, n)
        recs.append(dict(rows=["".join(rng.choice(list("acgu"), n))], p_left=np.sqrt(a), p_right=np.sqrt(b),
                         p_unpair=np.sqrt(np.maximum(0, 1 - a - b))))
    e = bpla.BplaSet(recs)
    xi, yi = np.divmod(np.arange(len(e) ** 2), len(e))
    p = bpla.make_params()
    check(bpla.gradients(ctx, p, e, e, xi, yi), O.bpla_gradients(p, e, e, xi, yi))
    # 150-300 nt records, x and y sets differ
    a = bpla.BplaSet([dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(3, 4, offset=50)])
    b = bpla.BplaSet([dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(3, 3, offset=90)])
    xi, yi = np.divmod(np.arange(12), 3)
    check(bpla.gradients(ctx, p, a, b, xi, yi), O.bpla_gradients(p, a, b, xi, yi))
    # the optimizer's matrices (bpla_optimizer.cpp:55-123): symmetric, diagonal = the pair (i, i)
    km, gm = bpla.gradient_matrices(ctx, p, e)
    wv, wg = O.bpla_gradients(p, e, e, np.arange(len(e)), np.arange(len(e)))
    assert np.array_equal(km, km.T) and np.array_equal(gm, gm.transpose(0, 2, 1)) and relerr(np.diag(km), wv) < TOL
    assert relerr(gm[0].diagonal(), wg[:, 0]) < TOL
    with pytest.raises(api.StemkError, match="no_bp = sw = 0"):
        bpla.gradients(ctx, bpla.make_params(sw=True), e, e, [0], [0])
