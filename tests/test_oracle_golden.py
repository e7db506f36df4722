"""The oracle (oracle/stemk_oracle.c) against the golden vectors generated from the unmodified reference, and,
where oracle/_ref is present, against the live reference.  Bit-exact: the restatement keeps the reference's
statement and operand order."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, TH
from oracle import oraclebind as O
from oracle import refbind as R
from stem_kernel_b200 import _lib as L
from stem_kernel_b200 import hostlib, synth


def oparams(kind, **kw):
    return O.Params.from_buffer_copy(L.make_params(kind, **kw))


@pytest.mark.parametrize("kind", range(9))
@pytest.mark.parametrize("band", [10, 0])
def test_gram_bit_exact(golden, kind, band):
    got = O.gram(oparams(kind, len_band=band), golden["flat"].desc(), False)
    assert np.array_equal(got, golden["z"][f"gram_k{kind}_b{band}"])


def test_normalised_gram_and_text(golden):
    z = golden["z"]
    got = O.gram(oparams(L.SU_STEM_STR), golden["flat"].desc(), True)
    want = z["gram_norm_k3_b10"]
    assert np.array_equal(got, want, equal_nan=True)
    # the empty-DAG record has k_stem = 0 but k_string > 0, so nothing is NaN here; LIBSVM text is byte-identical
    labels = [r["label"] for r in golden["recs"]]
    assert O.print_matrix(got, labels) == str(z["gram_norm_text_k3_b10"])


def test_empty_dag_gives_zero_and_nan_after_normalisation(golden):
    n = len(golden["recs"])
    empty = n - 2
    g = O.gram(oparams(L.SU_STEM), golden["flat"].desc(), False)
    assert np.all(g[empty] == 0.0) and np.all(g[:, empty] == 0.0)
    gn = O.gram(oparams(L.SU_STEM), golden["flat"].desc(), True)
    assert np.isnan(np.delete(gn[empty], empty)).all()      # 0/sqrt(0*x)  (SURVEY 8(b) error conventions)
    assert gn[empty, empty] == 1.0                             # kernel_matrix.cpp:570


def test_cross_row_diag(golden):
    z, md = golden["z"], golden["md"]
    test, train = hostlib.SeqSet(md[:5]), hostlib.SeqSet(md[5:])
    p = oparams(L.SU_STEM_STR)
    m, selfv = O.cross(p, test.desc(), train.desc())
    assert np.array_equal(m, z["cross_k3"]) and np.array_equal(selfv, z["cross_self_k3"])
    m, _ = O.cross(p, test.desc(), train.desc(), normalize=True)
    assert np.array_equal(m, z["cross_norm_k3"], equal_nan=True)
    one = hostlib.SeqSet([md[2]])
    sv = z["row_sv_index"]
    row, s = O.cross(p, one.desc(), train.desc(), sv_index=sv, init=-1.0)
    assert np.array_equal(row[0], z["row_sv_k3"]) and s[0] == float(z["row_sv_self_k3"])
    assert np.array_equal(O.diag(p, train.desc()), z["diag_k3"])
    assert np.array_equal(O.diag(p, train.desc(), sv_index=sv, init=-1.0), z["diag_sv_k3"])


def test_naive_string_kernel_golden():
    z = np.load(os.path.join(GOLDEN, "golden_naive.npz"))
    seqs = json.loads(str(z["seqs_json"]))
    flat = hostlib.SeqSet([hostlib.MData.seq_only([s]) for s in seqs])
    for gi, g in enumerate(z["gaps"]):
        got = O.gram(oparams(L.STR_NAIVE, gap=float(g)), flat.desc(), False)
        assert np.array_equal(got, z[f"gram_g{gi}"])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_bit_exact_on_fresh_inputs():
    recs = synth.make_config(1, 5, offset=100) + synth.make_config(3, 2, offset=50) + \
        [synth.alignment_like(11, i, n_rows=2 + i) for i in range(3)]
    md = [hostlib.MData.from_record(r, TH) for r in recs]
    ref = [R.RefMData.build(r["rows"], r["bp"], TH) for r in recs]
    flat = hostlib.SeqSet(md)   # owns the arrays the descriptor points into
    d = flat.desc()
    for kind in (L.SI_STEM, L.SU_STEM_STR, L.LSU_STEM_STR, L.STR_SIMPLE):
        want = R.RefKernel(kind, loop_gap=0.35, beta=0.5, gap=0.7, alpha=0.4, len_band=4).gram(ref)[0]
        got = O.gram(oparams(kind, loop_gap=0.35, beta=0.5, gap=0.7, alpha=0.4, len_band=4), d, False)
        assert np.array_equal(got, want)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_threaded_reference_equals_single_thread():
    recs = synth.make_config(1, 8)
    ref = [R.RefMData.build(r["rows"], r["bp"], TH) for r in recs]
    k = R.RefKernel(R.SU_STEM_STR)
    assert np.array_equal(k.gram(ref, n_th=1)[0], k.gram(ref, n_th=4)[0])


def test_pair_cost_model_matches_survey_formula(golden):
    """flops_stem = 2*U_match + 3*U_bf + 3*U_skip, checked on the single-hairpin record by hand."""
    n = len(golden["recs"])
    hp = n - 1
    d = golden["flat"].desc()
    ex = golden["md"][hp].export()
    V, E = len(ex["first"]), len(ex["edge_to"])
    nonleaf = [u for u in range(V) if ex["edge_off"][u + 1] > ex["edge_off"][u]]
    deg = {u: int(ex["edge_off"][u + 1] - ex["edge_off"][u]) for u in nonleaf}
    um = sum(deg[a] * deg[b] for a in nonleaf for b in nonleaf)   # band 0: every non-leaf pair
    ub = len(nonleaf) ** 2                                         # one (a,b) entry per node
    cells, flops = O.pair_cost(oparams(L.SU_STEM, len_band=0), d, hp, d, hp)
    assert cells == V * V and flops == 2 * um + 3 * ub + 3 * (2 * V * E)
