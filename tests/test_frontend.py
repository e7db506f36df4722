"""Host front end (stem_kernel_b200/host/frontend.cpp: rows + base-pair lists -> MData) against the reference
constructor's own MData (stem_kernel_lite/data.cpp:324-345), dumped into tests/golden/golden_mdata.npz by
tests/make_golden.py.  Bit-identical: node order, edges, gaps, float weights, base-pair profiles, roots, max_pa."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, TH
from oracle import refbind as R
from stem_kernel_b200 import hostlib, synth

KEYS = ["first", "last", "weight", "edge_off", "edge_to", "edge_gaps", "edge_w", "bpf_off", "bpf_a", "bpf_b", "bpf_f",
        "root", "max_pa", "profile", "seq_weight"]


def test_mdata_equals_reference_dump(golden):
    z = np.load(os.path.join(GOLDEN, "golden_mdata.npz"))
    for i, m in enumerate(golden["md"]):
        ex = m.export()
        for k in KEYS:
            assert np.array_equal(np.asarray(ex[k]), z[f"r{i}_{k}"]), (i, k)
        assert ex["n_seqs"] == float(z[f"r{i}_n_seqs"])


def test_children_before_parents_and_gaps(golden):
    for m in golden["md"]:
        ex = m.export()
        for u in range(len(ex["first"])):
            for e in range(ex["edge_off"][u], ex["edge_off"][u + 1]):
                c = ex["edge_to"][e]
                assert c < u                                     # data.cpp:193-244 post-order
                lp, lc = ex["last"][u] - ex["first"][u], ex["last"][c] - ex["first"][c]
                want = lp - 1 if lc == 0 else lp - lc - 2        # dag.h:25-26,32
                assert ex["edge_gaps"][e] == want


def test_empty_and_sequence_only_records(golden):
    n = len(golden["md"])
    s = golden["md"][n - 2].sizes()
    assert s["n_nodes"] == 0 and s["n_roots"] == 0 and s["n_weights"] == s["length"] == 20
    so = hostlib.MData.seq_only(["acgu-nry"])
    ex = so.export()
    assert so.sizes()["n_weights"] == 0 and so.sizes()["n_nodes"] == 0
    assert np.array_equal(ex["profile"][4], [0, 0, 0, 0, 1])           # '-' is a gap column
    assert np.array_equal(ex["profile"][5], [0.25, 0.25, 0.25, 0.25, 0])  # 'n' spread over the four bases
    assert np.array_equal(ex["profile"][6], [0.5, 0, 0.5, 0, 0])         # 'r' = a|g  (profile.cpp:11-29)


def test_errors():
    with pytest.raises(ValueError):
        hostlib.MData.seq_only(["acgu", "acg"])      # ragged alignment ("wrong alignment", data.cpp:574-578)
    z = np.zeros(0, dtype=np.int64)
    with pytest.raises(ValueError):                  # a pair closer than 2 columns cannot be represented
        hostlib.MData.build(["acgu"], [(np.array([1]), np.array([2]), np.array([0.5]))], TH)
    hostlib.MData.build(["acgu"], [(z, z, np.zeros(0))], TH)


def test_build_many_equals_one_by_one():
    recs = synth.make_config(3, 6)
    a = hostlib.build_many(recs, TH, n_threads=3)
    for r, m in zip(recs, a):
        one = hostlib.MData.from_record(r, TH).export()
        ex = m.export()
        for k in KEYS:
            assert np.array_equal(ex[k], one[k])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_constructor_on_fresh_inputs():
    recs = synth.make_config(3, 3, offset=500) + [synth.alignment_like(23, i, n_rows=5, L=80) for i in range(3)]
    for r in recs:
        ours = hostlib.MData.from_record(r, TH).export()
        ref = R.RefMData.build(r["rows"], r["bp"], TH).dump()
        for k in KEYS:
            assert np.array_equal(np.asarray(ours[k]), np.asarray(ref[k])), k
