"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref (run in the build
container, where /root/reference exists:  python tests/make_golden.py).  The reference ships no golden vectors
of its own, so these files pin the oracle and the CUDA path to the reference's own outputs:

  golden_pairs.npz     inputs (rows + sparse base-pair lists) of 17 records and, for every lite kernel class
                       x {len_band 10, 0}: the reference's Gram matrix; normalised Gram, rectangular
                       test x train matrix, sv_index row, diagonal and the printed text for SuStemStrKernel
  golden_mdata.npz     the reference constructor's MData (DAG, profiles, weights) for each record
  golden_naive.npz     string_kernel/ (naive) Gram on 8 raw strings, gap parsed as float
  golden_svm.npz       40-record C1 Gram (reference) and the vendored LIBSVM's 5-fold CV targets on it
  golden_nstem.npz     stem_kernel/ (the naive O(L^4) stem kernel, full_dp, banded partial_dp, partial_dp under pair-HMM windows): 9 short sequences, canonical pairs with and
                       without g-u, probability tables, a non-default parameter set; upper-triangle values
  golden_bpla.npz      bpla_kernel/ (BPLA / local-alignment kernels): 12 records (single sequences, alignments,
                       IUPAC, gaps, a length-1 and a 70-column record), their base-pairing profiles, and the
                       reference's upper-triangle values for {BP, noBP} x {sum form, Smith-Waterman}, and
                       compute_gradients (value + four derivatives per pair) for two parameter sets
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refbind as R  # noqa: E402
from stem_kernel_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TH = 0.01


def golden_records():
    recs = synth.make_config(1, 6) + synth.make_config(3, 3)
    recs += [synth.alignment_like(7, i, n_rows=(i % 4) + 1) for i in range(6)]
    z = np.zeros(0, dtype=np.int64)
    # no pair reaches the threshold: empty DAG, k_stem = 0 (stem_kernel.cpp:88-93 sums over no roots)
    recs.append(dict(rows=["acguacguacguacguacgu"], bp=[(z, z, np.zeros(0))], label=1))
    # a single isolated hairpin: one stem of 3 stacked pairs
    recs.append(dict(rows=["gggaaaaccc"], bp=[(np.array([1, 2, 3]), np.array([10, 9, 8]), np.array([0.9, 0.8, 0.7]))],
                     label=-1))
    return recs


def bpla_records():
    """Records of the BPLA golden file: C1-like single sequences with profiles from their base-pair lists, alignments
    with IUPAC codes and gaps and hand-made profiles, a length-1 record and a 70-column record (three 32-column chunks)."""
    from stem_kernel_b200 import bpla
    recs = [dict(rows=r["rows"], bp=r["bp"]) for r in synth.make_config(1, 5)]
    rng = np.random.default_rng(20260018)
    for rows in (["GGGAAACCC--A", "GCGAANCCCUUA"], ["acgu-ryacgu", "acguaNNacgu", "ac-uacgacgu"], ["a"], ["".join(rng.choice(list("acgu"), 70))],
                 ["gggg", "cccc"], ["u" * 33], ["acgu" * 8 + "n"]):
        n = len(rows[0])
        a, b = rng.uniform(0, 0.6, n), rng.uniform(0, 0.4, n)
        recs.append(dict(rows=rows, p_left=np.sqrt(a).astype(np.float32), p_right=np.sqrt(b).astype(np.float32),
                         p_unpair=np.sqrt(np.maximum(0, 1 - a - b)).astype(np.float32)))
    return recs


def nstem_inputs():
    """Sequences (lower case) and probability tables of the naive-stem-kernel golden file."""
    from stem_kernel_b200 import nstem
    recs = synth.make_config(1, 3)
    seqs = [r["rows"][0].lower()[:34] for r in recs] + ["gggaaaccc", "a", "gcgcuuuugcgc", "gggggaaaauuuuuccccc", "acgu" * 9, "au"]
    rng = np.random.default_rng(20260019)
    tabs = [nstem.dense_bp(len(r["rows"][0]), r["bp"][0])[:34, :34].copy() for r in recs]
    for s in seqs[3:]:
        n = len(s)
        tabs.append(np.triu(rng.uniform(0, 1, (n, n)) * (rng.uniform(0, 1, (n, n)) < 0.15), 3).astype(np.float32))
    return seqs, tabs


def pack_records(recs):
    rows = json.dumps([r["rows"] for r in recs])
    labels = np.array([r["label"] for r in recs], dtype=np.int32)
    bi, bj, bp, off = [], [], [], [0]
    for r in recs:
        for (a, b, p) in r["bp"]:
            bi.append(np.asarray(a, dtype=np.uint32)); bj.append(np.asarray(b, dtype=np.uint32))
            bp.append(np.asarray(p, dtype=np.float64)); off.append(off[-1] + len(a))
    return dict(rows_json=np.array(rows), labels=labels, bp_i=np.concatenate(bi), bp_j=np.concatenate(bj),
                bp_p=np.concatenate(bp), bp_off=np.array(off, dtype=np.int64))


def main():
    os.makedirs(OUT, exist_ok=True)
    recs = golden_records()
    ref = [R.RefMData.build(r["rows"], r["bp"], TH) for r in recs]
    out = pack_records(recs)
    for kind in range(9):
        for band in (10, 0):
            k = R.RefKernel(kind, len_band=band)
            out[f"gram_k{kind}_b{band}"] = k.gram(ref)[0]
    k = R.RefKernel(R.SU_STEM_STR)
    labels = [r["label"] for r in recs]
    gn, _, text = k.gram(ref, normalize=True, labels=labels, want_text=True)
    out["gram_norm_k3_b10"] = gn
    out["gram_norm_text_k3_b10"] = np.array(text)
    test, train = ref[:5], ref[5:]
    m, selfv, _ = k.cross(test, train, norm_test=True, normalize=False)
    out["cross_k3"] = m; out["cross_self_k3"] = selfv
    m, selfv, _ = k.cross(test, train, norm_test=True, normalize=True)
    out["cross_norm_k3"] = m
    sv = np.array([1, 4, 7], dtype=np.uint32)
    row, s, _ = k.row(ref[2], train, sv_index=sv, init=-1.0)
    out["row_sv_index"] = sv; out["row_sv_k3"] = row; out["row_sv_self_k3"] = np.array(s)
    out["diag_k3"] = k.diag(train)[0]
    out["diag_sv_k3"] = k.diag(train, sv_index=sv, init=-1.0)[0]
    np.savez_compressed(os.path.join(OUT, "golden_pairs.npz"), **out)

    md = {}
    for i, d in enumerate(ref):
        for key, v in d.dump().items():
            md[f"r{i}_{key}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "golden_mdata.npz"), **md)

    seqs = [r["rows"][0] for r in synth.make_config(2, 6)] + ["ACGUacgu", "gattaca"]
    gaps = [np.float32(0.8), np.float32(1.0)]
    nv = dict(seqs_json=np.array(json.dumps(seqs)), gaps=np.array(gaps, dtype=np.float32))
    for gi, g in enumerate(gaps):
        nv[f"gram_g{gi}"] = R.naive_gram(float(g), seqs)[0]
    np.savez_compressed(os.path.join(OUT, "golden_naive.npz"), **nv)

    recs1 = synth.make_config(1, 40)
    ref1 = [R.RefMData.build(r["rows"], r["bp"], TH) for r in recs1]
    K = R.RefKernel(R.SU_STEM_STR).gram(ref1, normalize=True)[0]
    y = np.array([r["label"] for r in recs1], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "golden_svm.npz"), gram=K, y=y, cv_target=R.svm_cv(K, y, 1.0, 5, 1))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

    # ---- BPLA / local-alignment kernels (bpla_kernel/bpla_kernel.cpp via oracle/ref_harness_bpla.cpp)
    from stem_kernel_b200 import bpla  # noqa: E402
    brecs = bpla_records()
    bs = bpla.BplaSet(brecs)
    xi, yi = np.triu_indices(len(bs))
    g = dict(rows_json=json.dumps([r["rows"] for r in brecs]), col_off=bs.col_off, p_left=bs.p_left, p_right=bs.p_right,
             p_unpair=bs.p_unpair, xi=xi.astype(np.uint32), yi=yi.astype(np.uint32))
    for no_bp in (0, 1):
        for sw in (0, 1):
            g[f"k_nobp{no_bp}_sw{sw}"] = R.bpla_pairs(bpla.make_params(no_bp=no_bp, sw=sw), bs, bs, xi, yi)
    g["k_custom"] = R.bpla_pairs(bpla.make_params(gap=-3.0, ext=-0.25, alpha=2.0, beta=0.3, score=np.arange(16.0).reshape(4, 4) / 4 - 1),
                                 bs, bs, xi, yi)
    # BPLAKernel::compute_gradients (bpla_kernel.cpp:387-402): values and d/d{alpha, beta, gap, ext}, default and custom parameters
    g["grad_value"], g["grad"] = R.bpla_gradients(bpla.make_params(), bs, bs, xi, yi)
    g["grad_value_custom"], g["grad_custom"] = R.bpla_gradients(
        bpla.make_params(gap=-3.0, ext=-0.25, alpha=2.0, beta=0.3, score=np.arange(16.0).reshape(4, 4) / 4 - 1), bs, bs, xi, yi)
    np.savez_compressed(os.path.join(OUT, "golden_bpla.npz"), **g)

    # ---- naive stem kernel (stem_kernel/stem_kernel.cpp via oracle/ref_harness_nstem.cpp)
    from stem_kernel_b200 import nstem  # noqa: E402
    seqs, tabs = nstem_inputs()
    sa, sb = nstem.NstemSet(seqs), nstem.NstemSet(seqs, tabs)
    xi, yi = np.triu_indices(len(seqs))
    n = dict(seqs_json=json.dumps(seqs), tables=np.concatenate([t.reshape(-1) for t in tabs]), xi=xi.astype(np.uint32),
             yi=yi.astype(np.uint32))
    n["k_normal"] = R.nstem_pairs(nstem.make_params(), sa, sa, xi, yi)
    n["k_wobble"] = R.nstem_pairs(nstem.make_params(use_gu=True), sa, sa, xi, yi)
    n["k_table"] = R.nstem_pairs(nstem.make_params(bp_mode=1, bp_bound=0.05), sb, sb, xi, yi)
    n["k_custom"] = R.nstem_pairs(nstem.make_params(loop=1, gap=0.6, stack=1.7, subst=0.3), sa, sa, xi, yi)
    n["k_default_bound"] = R.nstem_pairs(nstem.make_params(bp_bound=1.0), sa, sa, xi, yi)
    # the banded partial_dp (stem_kernel.cpp:113-280, band-only constraints :77-83)
    for band in (3, 8):
        n[f"k_normal_band{band}"] = R.nstem_pairs(nstem.make_params(), sa, sa, xi, yi, band=band)
        n[f"k_table_band{band}"] = R.nstem_pairs(nstem.make_params(bp_mode=1, bp_bound=0.05), sb, sb, xi, yi, band=band)
    # partial_dp under the pair-HMM constraints of ali_bound > 0 (alignment_constraints, stem_kernel.cpp:14-67; phmm.cpp),
    # alone and narrowed by a band (:57-66): the reference's own windows (c_low / c_high per row) and its values
    for tag, band, ab in (("ali", 0, 0.3), ("ali_band4", 4, 0.3)):
        wins = [R.nstem_windows(seqs[a], seqs[b], band=band, ali_bound=ab) for a, b in zip(xi, yi)]
        n[f"win_low_{tag}"] = np.concatenate([w[0] for w in wins])
        n[f"win_high_{tag}"] = np.concatenate([w[1] for w in wins])
        n[f"k_normal_{tag}"] = R.nstem_pairs(nstem.make_params(), sa, sa, xi, yi, band=band, ali_bound=ab)
        n[f"k_table_{tag}"] = R.nstem_pairs(nstem.make_params(bp_mode=1, bp_bound=0.05), sb, sb, xi, yi, band=band, ali_bound=ab)
    np.savez_compressed(os.path.join(OUT, "golden_nstem.npz"), **n)
