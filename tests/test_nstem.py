"""Naive stem kernel (stem_kernel/stem_kernel.cpp:282-351 full_dp, base-pair classes :353-420).

CPU: the plain-C restatement against the golden values of the compiled reference, bit for bit, and against the
reference itself on fresh inputs where oracle/_ref exists.  GPU: the CUDA kernel through the C ABI
(stemk_nstem_pairs) against the oracle to 1e-9 relative, on the golden inputs, argument order swapped (the kernel
stages the shorter sequence), rectangular x / y sets, a 60-nt pair, and the degenerate default bound."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, need_gpu, relerr
from oracle import oraclebind as O
from oracle import refbind as R
from stem_kernel_b200 import nstem, synth

TOL = 1e-9


def golden_sets():
    z = np.load(os.path.join(GOLDEN, "golden_nstem.npz"))
    seqs = json.loads(str(z["seqs_json"]))
    tabs, at = [], 0
    for s in seqs:
        n = len(s)
        tabs.append(z["tables"][at:at + n * n].reshape(n, n))
        at += n * n
    return nstem.NstemSet(seqs), nstem.NstemSet(seqs, tabs), z


CASES = [("k_normal", False, dict()), ("k_wobble", False, dict(use_gu=True)), ("k_table", True, dict(bp_mode=1, bp_bound=0.05)),
         ("k_custom", False, dict(loop=1, gap=0.6, stack=1.7, subst=0.3)), ("k_default_bound", False, dict(bp_bound=1.0))]


def test_restatement_is_bit_identical_to_the_reference_golden():
    sa, sb, z = golden_sets()
    for key, tab, kw in CASES:
        s = sb if tab else sa
        assert np.array_equal(O.nstem_pairs(nstem.make_params(**kw), s, s, z["xi"], z["yi"]), z[key]), key
    assert np.all(z["k_default_bound"] == 1.0)          # the program's own default bound: no pair ever counts
    assert z["k_normal"].min() >= 1.0 and z["k_normal"].max() > 1e3


@pytest.mark.skipif(not os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_nstem.so")), reason="oracle/_ref not built")
def test_restatement_matches_the_compiled_reference_on_fresh_inputs():
    rng = np.random.default_rng(11)
    seqs = ["".join(rng.choice(list("acgu"), n)) for n in (1, 5, 12, 20, 27)] + ["gggcaaagccc"]
    s = nstem.NstemSet(seqs)
    xi, yi = np.divmod(np.arange(len(seqs) ** 2), len(seqs))
    for kw in (dict(), dict(use_gu=True, loop=0), dict(gap=0.3, stack=2.0, subst=0.9)):
        p = nstem.make_params(**kw)
        assert np.array_equal(O.nstem_pairs(p, s, s, xi, yi), R.nstem_pairs(p, s, s, xi, yi))


def test_banded_restatement_is_bit_identical_to_the_reference():
    """StemKernel::partial_dp with the band-only constraints (stem_kernel.cpp:14-83, 113-280): golden values of the
    compiled reference, and, where oracle/_ref exists, fresh inputs incl. a band wider than the sequences (= full_dp).
    (Oracle only so far: the CUDA kernel covers full_dp)"""
    sa, sb, z = golden_sets()
    for band in (3, 8):
        assert np.array_equal(O.nstem_pairs_banded(nstem.make_params(), band, sa, sa, z["xi"], z["yi"]), z[f"k_normal_band{band}"])
        assert np.array_equal(O.nstem_pairs_banded(nstem.make_params(bp_mode=1, bp_bound=0.05), band, sb, sb, z["xi"], z["yi"]),
                              z[f"k_table_band{band}"])
    assert not np.array_equal(z["k_normal_band3"], z["k_normal"])
    if os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_nstem.so")):
        rng = np.random.default_rng(12)
        seqs = ["".join(rng.choice(list("acgu"), n)) for n in (2, 9, 16, 23)] + ["gggcaaagccc"]
        s = nstem.NstemSet(seqs)
        xi, yi = np.divmod(np.arange(len(seqs) ** 2), len(seqs))
        p = nstem.make_params(use_gu=True, loop=1)
        for band in (1, 4, 30):
            assert np.array_equal(O.nstem_pairs_banded(p, band, s, s, xi, yi), R.nstem_pairs(p, s, s, xi, yi, band=band))
        assert np.array_equal(O.nstem_pairs_banded(p, 30, s, s, xi, yi), O.nstem_pairs(p, s, s, xi, yi))


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_golden_swapped_and_rectangular():
    need_gpu()
    from stem_kernel_b200 import api, _lib as L
    ctx = api.Context(L.make_params(L.STR_SIMPLE))
    sa, sb, z = golden_sets()
    for key, tab, kw in CASES:
        s, p = (sb if tab else sa), nstem.make_params(**kw)
        assert relerr(nstem.pairs(ctx, p, s, s, z["xi"], z["yi"]), z[key]) < TOL, key
        assert relerr(nstem.pairs(ctx, p, s, s, z["yi"], z["xi"]), O.nstem_pairs(p, s, s, z["yi"], z["xi"])) < TOL, key
    rng = np.random.default_rng(3)
    a = nstem.NstemSet(["".join(rng.choice(list("acgu"), n)) for n in (60, 33, 7)])
    b = nstem.NstemSet(["".join(rng.choice(list("acgu"), n)) for n in (58, 1, 40, 12)])
    xi, yi = np.divmod(np.arange(12), 4)
    p = nstem.make_params(use_gu=True)
    assert relerr(nstem.pairs(ctx, p, a, b, xi, yi), O.nstem_pairs(p, a, b, xi, yi)) < TOL
    with pytest.raises(api.StemkError, match="out of range"):
        nstem.pairs(ctx, p, a, b, [3], [0])
    with pytest.raises(api.StemkError, match="tables missing"):
        nstem.pairs(ctx, nstem.make_params(bp_mode=1), a, b, [0], [0])


@pytest.mark.gpu
def test_gpu_config1_like_probability_tables():
    """C1-like records (~75 nt) with their thresholded base-pair probabilities as tables: 8 records, 36 pairs."""
    need_gpu()
    from stem_kernel_b200 import api, _lib as L
    ctx = api.Context(L.make_params(L.STR_SIMPLE))
    recs = synth.make_config(1, 8, offset=40)
    seqs = [r["rows"][0].lower() for r in recs]
    s = nstem.NstemSet(seqs, [nstem.dense_bp(len(q), r["bp"][0], th=0.01) for q, r in zip(seqs, recs)])
    xi, yi = np.triu_indices(len(s))
    p = nstem.make_params(bp_mode=1, bp_bound=0.01)
    got, want = nstem.pairs(ctx, p, s, s, xi, yi), O.nstem_pairs(p, s, s, xi, yi)
    assert relerr(got, want) < TOL and want.max() > 1.0


@pytest.mark.gpu
def test_gpu_banded_partial_dp():
    """stemk_nstem_pairs_banded (partial_dp with the band-only constraints) against the reference's golden values and
    the oracle: canonical pairs and probability tables, bands 1 ... wider than the sequences (= full_dp), x and y
    sets of different lengths in both roles."""
    need_gpu()
    from stem_kernel_b200 import api, _lib as L
    ctx = api.Context(L.make_params(L.STR_SIMPLE))
    sa, sb, z = golden_sets()
    errs = {}
    for band in (3, 8):
        errs[("normal", band)] = relerr(nstem.pairs_banded(ctx, nstem.make_params(), band, sa, sa, z["xi"], z["yi"]), z[f"k_normal_band{band}"])
        errs[("table", band)] = relerr(nstem.pairs_banded(ctx, nstem.make_params(bp_mode=1, bp_bound=0.05), band, sb, sb, z["xi"], z["yi"]),
                                       z[f"k_table_band{band}"])
    rng = np.random.default_rng(13)
    a = nstem.NstemSet(["".join(rng.choice(list("acgu"), n)) for n in (1, 7, 19, 33, 48)] + ["gggcaaagccc"])
    b = nstem.NstemSet(["".join(rng.choice(list("acgu"), n)) for n in (2, 11, 26, 40)])
    xi, yi = np.divmod(np.arange(len(a.seqs) * len(b.seqs)), len(b.seqs))
    p = nstem.make_params(use_gu=True, loop=1, gap=0.7)
    for band in (1, 5, 12, 60):
        errs[("ab", band)] = relerr(nstem.pairs_banded(ctx, p, band, a, b, xi, yi), O.nstem_pairs_banded(p, band, a, b, xi, yi))
        errs[("ba", band)] = relerr(nstem.pairs_banded(ctx, p, band, b, a, yi, xi), O.nstem_pairs_banded(p, band, b, a, yi, xi))
    errs[("full", 60)] = relerr(nstem.pairs_banded(ctx, p, 60, a, b, xi, yi), O.nstem_pairs(p, a, b, xi, yi))
    assert max(errs.values()) < TOL, errs
    with pytest.raises(api.StemkError, match="band must be positive"):
        nstem.pairs_banded(ctx, p, 0, a, b, xi, yi)


def golden_windows(z, tag, sa):
    """The reference's own c_low / c_high (alignment_constraints, stem_kernel.cpp:14-67) per golden pair."""
    wins, at = [], 0
    for a in z["xi"]:
        n = len(sa.seqs[int(a)]) + 1
        wins.append((z[f"win_low_{tag}"][at:at + n], z[f"win_high_{tag}"][at:at + n]))
        at += n
    return wins


def test_window_restatement_is_bit_identical_to_the_reference_with_pair_hmm_constraints():
    """partial_dp under the constraints of ali_bound > 0 (pair-HMM posteriors, phmm.cpp; alone and narrowed by a band,
    stem_kernel.cpp:25-67): the restatement fed the reference's own windows gives the reference's values bit for bit."""
    sa, sb, z = golden_sets()
    for tag in ("ali", "ali_band4"):
        wins = golden_windows(z, tag, sa)
        assert max(int((h - l).max()) for l, h in wins) > 8          # wider than any band tested: unaligned stretches
        assert np.array_equal(O.nstem_pairs_windows(nstem.make_params(), sa, sa, z["xi"], z["yi"], wins), z[f"k_normal_{tag}"])
        assert np.array_equal(O.nstem_pairs_windows(nstem.make_params(bp_mode=1, bp_bound=0.05), sb, sb, z["xi"], z["yi"], wins),
                              z[f"k_table_{tag}"])
    # the band-only windows through the same entry give the banded values
    band = 3
    wins = []
    for a, b in zip(z["xi"], z["yi"]):
        lx, ly = len(sa.seqs[int(a)]), len(sa.seqs[int(b)])
        c = (np.arange(lx + 1) / lx * ly + 0.5).astype(np.uint32)
        wins.append((np.where(c < band, 0, c - band), np.minimum(c + band, ly)))
    assert np.array_equal(O.nstem_pairs_windows(nstem.make_params(), sa, sa, z["xi"], z["yi"], wins), z["k_normal_band3"])


@pytest.mark.skipif(not os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_nstem.so")), reason="oracle/_ref not built")
def test_window_restatement_matches_the_compiled_reference_on_fresh_inputs():
    rng = np.random.default_rng(23)
    seqs = ["".join(rng.choice(list("acgu"), n)) for n in (9, 14, 21, 26)] + ["gggcaaagccc", "gggcuaaagcccu"]
    s = nstem.NstemSet(seqs)
    xi, yi = np.divmod(np.arange(len(seqs) ** 2), len(seqs))
    for band, ab in ((0, 0.2), (2, 0.5), (5, 0.05)):
        wins = [R.nstem_windows(seqs[a], seqs[b], band=band, ali_bound=ab) for a, b in zip(xi, yi)]
        p = nstem.make_params(use_gu=True)
        assert np.array_equal(O.nstem_pairs_windows(p, s, s, xi, yi, wins), R.nstem_pairs(p, s, s, xi, yi, band=band, ali_bound=ab))


@pytest.mark.gpu
def test_gpu_windows_match_the_reference_golden_and_the_oracle():
    """stemk_nstem_pairs_windows: the device kernel under the reference's pair-HMM windows."""
    need_gpu()
    from stem_kernel_b200 import _lib as L
    from stem_kernel_b200 import api
    sa, sb, z = golden_sets()
    ctx = api.Context(L.make_params(L.STR_SIMPLE), device=0)
    for tag in ("ali", "ali_band4"):
        wins = golden_windows(z, tag, sa)
        got = nstem.pairs_windows(ctx, nstem.make_params(), sa, sa, z["xi"], z["yi"], wins)
        assert relerr(got, z[f"k_normal_{tag}"]) < TOL
        got = nstem.pairs_windows(ctx, nstem.make_params(bp_mode=1, bp_bound=0.05), sb, sb, z["xi"], z["yi"], wins)
        assert relerr(got, z[f"k_table_{tag}"]) < TOL
    # fresh sequences, arbitrary monotone windows (not produced by any band): against the oracle
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("acgu"), n)) for n in (11, 17, 24, 30)]
    s = nstem.NstemSet(seqs)
    xi, yi = np.divmod(np.arange(len(seqs) ** 2), len(seqs))
    wins = []
    for a, b in zip(xi, yi):
        lx, ly = len(seqs[a]), len(seqs[b])
        c = (np.arange(lx + 1) / lx * ly + 0.5).astype(np.int64)
        w = rng.integers(1, 7, size=lx + 1)
        lo = np.maximum.accumulate(np.clip(c - w, 0, ly))
        hi = np.maximum.accumulate(np.maximum(np.clip(c + w[::-1], 0, ly), lo))
        wins.append((lo.astype(np.uint32), hi.astype(np.uint32)))
    p = nstem.make_params(use_gu=True, gap=0.6)
    assert relerr(nstem.pairs_windows(ctx, p, s, s, xi, yi, wins), O.nstem_pairs_windows(p, s, s, xi, yi, wins)) < TOL
    # malformed windows are refused
    bad = [(w[0][:-1], w[1][:-1]) for w in wins]
    with pytest.raises(api.StemkError):
        nstem.pairs_windows(ctx, p, s, s, xi, yi, bad)
