"""Parity of the CUDA path, called through the C ABI (include/stemk.h via stem_kernel_b200/api.py), against
 (1) the committed golden vectors generated from the unmodified reference,
 (2) the oracle on freshly seeded inputs, edge cases included,
 (3) size-independent properties at BASELINE.json's full sizes.
Tolerance: 1e-9 relative in fp64 on every entry (north_star; SURVEY 8(d) parity gate); NaN/0 patterns identical.
The CUDA kernels reorder sums (K tables eliminated, warp-parallel sweeps), so bit-equality is not expected."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, TH, need_gpu, relerr
from oracle import oraclebind as O
from oracle import refbind as R
from stem_kernel_b200 import _lib as L
from stem_kernel_b200 import api, hostlib, synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def oparams(p):
    return O.Params.from_buffer_copy(p)


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    need_gpu()


# ------------------------------------------------------------------ (1) golden vectors of the reference
@pytest.mark.parametrize("kind", range(9))
@pytest.mark.parametrize("band", [10, 0])
@pytest.mark.parametrize("fast", ["1", "0"])
def test_golden_gram(golden, kind, band, fast):
    """fast=1: single-sequence records take the separable fast stem kernel, alignments the general one;
    fast=0 forces the general kernel for every pair (stemk_set_option(STEMK_OPT_FORCE_GENERAL))."""
    if fast == "0" and kind in (L.STR_SUBST, L.STR_SIMPLE, L.LSU_STR):
        pytest.skip("no stem part")
    ctx = api.Context(L.make_params(kind, len_band=band)).set_option(L.OPT_FORCE_GENERAL, fast == "0")
    got = ctx.gram(ctx.upload(golden["flat"]))
    assert relerr(got, golden["z"][f"gram_k{kind}_b{band}"]) < TOL
    assert np.array_equal(got, got.T)
    ctx.close()


@pytest.mark.parametrize("band", [10, 0])
def test_golden_gram_unstaged_kernel(golden, band):
    """The any-size variant of the general stem kernel (nothing staged in shared memory) on the golden records:
    STEMK_OPT_FORCE_UNSTAGED sends every pair of the general kernel to it."""
    ctx = api.Context(L.make_params(L.SU_STEM_STR, len_band=band))
    ctx.set_option(L.OPT_FORCE_GENERAL, True).set_option(L.OPT_FORCE_UNSTAGED, True)
    got = ctx.gram(ctx.upload(golden["flat"]))
    assert relerr(got, golden["z"][f"gram_k{L.SU_STEM_STR}_b{band}"]) < TOL
    ctx.close()


@pytest.mark.parametrize("band", [10, 0])
def test_records_of_any_size(band):
    """Records beyond every shared-memory limit (700-1500 nt: 1 200-2 400 non-leaf DAG nodes, 7 000-16 000 inner edges)
    next to ordinary ones and an alignment: the reference's operator() has no size limit; the
    classifier sends the pairs with a large record to the unstaged kernel, the others stay where they were."""
    recs = [synth.ncrna_like(777, i, 700, 900) for i in range(2)] + [synth.ncrna_like(778, 0, 1200, 1500)]
    recs += synth.make_config(3, 4, offset=50) + [synth.alignment_like(5, 1, n_rows=3, L=70)]
    flat = hostlib.SeqSet([hostlib.MData.from_record(r, TH) for r in recs])
    p = L.make_params(L.SU_STEM, len_band=band)
    ctx = api.Context(p)
    ds = ctx.upload(flat)
    want = O.gram(oparams(p), flat.desc(), False)
    assert relerr(ctx.gram(ds), want) < TOL
    # rectangular, large records on either side (the oracle's own rectangular matrix: the stem kernel is not symmetric
    # in its arguments and a row is kernel(train_j, test_i), kernel_matrix.cpp:164-171)
    smallf = hostlib.SeqSet([hostlib.MData.from_record(r, TH) for r in recs[3:]])
    small = ctx.upload(smallf)
    m, sv = ctx.cross(small, ds)
    wm, wsv = O.cross(oparams(p), smallf.desc(), flat.desc())
    assert relerr(m, wm) < TOL and relerr(sv, wsv) < TOL
    m, sv = ctx.cross(ds, small)
    wm, wsv = O.cross(oparams(p), flat.desc(), smallf.desc())
    assert relerr(m, wm) < TOL and relerr(sv, wsv) < TOL
    ctx.close()


def test_golden_normalised_text_is_byte_identical(golden):
    z = golden["z"]
    ctx = api.Context(L.make_params(L.SU_STEM_STR))
    got = ctx.gram(ctx.upload(golden["flat"]), normalize=True)
    assert relerr(got, z["gram_norm_k3_b10"]) < TOL
    labels = ["%+d" % r["label"] for r in golden["recs"]]
    assert api.format_matrix(got, labels) == str(z["gram_norm_text_k3_b10"])


def test_golden_cross_row_diag(golden):
    z, md = golden["z"], golden["md"]
    ctx = api.Context(L.make_params(L.SU_STEM_STR))
    test, train = ctx.upload(md[:5]), ctx.upload(md[5:])
    m, selfv = ctx.cross(test, train)
    assert relerr(m, z["cross_k3"]) < TOL and relerr(selfv, z["cross_self_k3"]) < TOL
    m, _ = ctx.cross(test, train, normalize=True)
    assert relerr(m, z["cross_norm_k3"]) < TOL
    sv = z["row_sv_index"]
    row, s = ctx.cross(ctx.upload([md[2]]), train, sv_index=sv, init=-1.0)
    assert relerr(row[0], z["row_sv_k3"]) < TOL and relerr(s, [float(z["row_sv_self_k3"])]) < TOL
    assert np.all(row[0][np.setdiff1d(np.arange(len(md) - 5), sv)] == -1.0)    # untouched outside sv_index
    assert relerr(ctx.diag(train), z["diag_k3"]) < TOL
    assert relerr(ctx.diag(train, sv_index=sv, init=-1.0), z["diag_sv_k3"]) < TOL


def test_golden_naive_string_kernel():
    z = np.load(os.path.join(GOLDEN, "golden_naive.npz"))
    seqs = json.loads(str(z["seqs_json"]))
    md = [hostlib.MData.seq_only([s]) for s in seqs]
    for gi, g in enumerate(z["gaps"]):
        k = api.NaiveStringKernel(gap=float(g))
        assert relerr(k.ctx.gram(k.ctx.upload(md)), z[f"gram_g{gi}"]) < TOL


def test_kernel_functor_single_pair(golden):
    """value_type operator()(const Data&, const Data&) -- the reference's per-pair interface."""
    k = api.SuStemStrKernel()
    want = golden["z"]["gram_k3_b10"]
    assert abs(k(golden["md"][0], golden["md"][7]) - want[0, 7]) <= TOL * abs(want[0, 7])


# ------------------------------------------------------------------ (2) oracle on fresh inputs, edge cases
def fresh_mixed(seed_off=0):
    recs = synth.make_config(1, 10, offset=300 + seed_off) + synth.make_config(3, 8, offset=40 + seed_off)
    recs += [synth.alignment_like(31 + seed_off, i, n_rows=(i % 5) + 1, L=50 + 7 * i) for i in range(8)]
    return recs


@pytest.mark.parametrize("kind", range(9))
def test_oracle_fresh_inputs_nondefault_params(kind):
    recs = fresh_mixed()
    flat = hostlib.SeqSet([hostlib.MData.from_record(r, TH) for r in recs])
    p = L.make_params(kind, loop_gap=0.37, beta=0.45, stack=1.7, covar=0.6, gap=0.66, alpha=0.33, match=1.2,
                      mismatch=0.7, len_band=(3 if kind % 2 else 0))
    ctx = api.Context(p)
    ds = ctx.upload(flat)
    for normalize in (False, True):
        assert relerr(ctx.gram(ds, normalize), O.gram(oparams(p), flat.desc(), normalize)) < TOL


@pytest.mark.parametrize("g", [0.9, 0.05, 0.001])
def test_loop_gap_range(g):
    """Large and tiny loop gaps: the separable fast path rescales rows by g^(+-len/2); when those powers leave its
    safe range (g = 0.001 on 300 nt) the records must fall back to the general kernel, not lose accuracy."""
    recs = synth.make_config(3, 6, offset=1300) + synth.make_config(1, 4, offset=1300)
    flat = hostlib.SeqSet([hostlib.MData.from_record(r, TH) for r in recs])
    for band in (10, 0):
        p = L.make_params(L.SU_STEM, loop_gap=g, len_band=band)
        ctx = api.Context(p)
        assert relerr(ctx.gram(ctx.upload(flat)), O.gram(oparams(p), flat.desc(), False)) < TOL


def test_fast_and_general_kernels_agree_and_fast_is_deterministic():
    recs = synth.make_config(3, 40, offset=2100)
    md = hostlib.build_many(recs, TH)
    p = L.make_params(L.SU_STEM)
    cg = api.Context(p).set_option(L.OPT_FORCE_GENERAL, 1)
    g_general = cg.gram(cg.upload(md))
    cf = api.Context(p)
    ds = cf.upload(md)
    g_fast = cf.gram(ds)
    assert relerr(g_fast, g_general) < 1e-11
    assert np.array_equal(g_fast, cf.gram(ds))          # fixed-order reduction: bit-reproducible
    perm = np.random.default_rng(0).permutation(40)
    gp = cf.gram(cf.upload([md[i] for i in perm]))
    a, b = np.triu_indices(40)
    kept = perm[a] <= perm[b]
    assert np.array_equal(gp[a[kept], b[kept]], g_fast[perm[a[kept]], perm[b[kept]]])


def test_threshold_changes_dag_density():
    """Denser DAGs (lower threshold keeps every background pair) and sparser ones (only planted stems)."""
    recs = synth.make_config(3, 6, offset=900)
    for th in (0.001, 0.2, 0.5):
        flat = hostlib.SeqSet([hostlib.MData.from_record(r, th) for r in recs])
        p = L.make_params(L.SU_STEM)
        ctx = api.Context(p)
        assert relerr(ctx.gram(ctx.upload(flat)), O.gram(oparams(p), flat.desc(), False)) < TOL


def test_empty_dag_zero_and_nan_pattern(golden):
    n = len(golden["recs"])
    empty = n - 2
    p = L.make_params(L.SU_STEM)
    ctx = api.Context(p)
    ds = ctx.upload(golden["flat"])
    g = ctx.gram(ds)
    assert np.all(g[empty] == 0.0) and np.all(g[:, empty] == 0.0)
    gn = ctx.gram(ds, normalize=True)
    want = O.gram(oparams(p), golden["flat"].desc(), True)
    assert np.array_equal(np.isnan(gn), np.isnan(want)) and gn[empty, empty] == 1.0
    assert relerr(gn, want) < TOL


def test_stem_kernel_on_sequence_only_records_is_zero():
    md = [hostlib.MData.seq_only([r["rows"][0]]) for r in synth.make_config(2, 4)]
    ctx = api.Context(L.make_params(L.SU_STEM))
    assert np.all(ctx.gram(ctx.upload(md)) == 0.0)
    p = L.make_params(L.SU_STEM_STR)
    ctx = api.Context(p)
    flat = hostlib.SeqSet(md)
    assert relerr(ctx.gram(ctx.upload(flat)), O.gram(oparams(p), flat.desc(), False)) < TOL


@pytest.mark.parametrize("kind", [L.STR_SUBST, L.STR_SIMPLE, L.STR_NAIVE])
def test_string_kernels_ragged_and_wide(kind):
    """Lengths 1..900: one-column records, lengths around the 32*CW tile edges, and multi-tile sweeps."""
    rng = np.random.default_rng(77)
    lens = [1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 383, 384, 385, 640, 900]
    seqs = ["".join(rng.choice(list("acgu"), size=n)) for n in lens]
    seqs[3] = seqs[3][:10] + "n-ry" + seqs[3][14:]            # IUPAC / gap columns take the general path
    flat = hostlib.SeqSet([hostlib.MData.seq_only([s]) for s in seqs])
    p = L.make_params(kind, gap=0.8 if kind != L.STR_NAIVE else float(np.float32(0.9)))
    ctx = api.Context(p)
    assert relerr(ctx.gram(ctx.upload(flat)), O.gram(oparams(p), flat.desc(), False)) < TOL


def test_weighted_and_unweighted_records_mix():
    """use_weight needs BOTH records to carry weights (string_kernel.cpp:77,93)."""
    recs = synth.make_config(1, 4)
    md = [hostlib.MData.from_record(r, TH) for r in recs[:2]] + [hostlib.MData.seq_only(r["rows"]) for r in recs[2:]]
    flat = hostlib.SeqSet(md)
    p = L.make_params(L.STR_SUBST)
    ctx = api.Context(p)
    assert relerr(ctx.gram(ctx.upload(flat)), O.gram(oparams(p), flat.desc(), False)) < TOL


def test_hand_made_dag_with_leaf_root_and_shared_children():
    """A DAG the front end never emits: a leaf that is also a root, a child shared by two parents, weights != 1."""
    f = dict(first=[2, 5, 1, 0, 9], last=[2, 8, 9, 11, 9], weight=[1.0, 0.7, 0.5, 0.9, 1.0],
             edge_off=[0, 0, 1, 3, 5, 5], edge_to=[0, 1, 0, 2, 1], edge_gaps=[2, 1, 7, 0, 5],
             edge_w=[1.0, 0.5, 1.0, 2.0, 1.0], bpf_off=[0, 0, 1, 3, 4, 4], bpf_a=[2, 0, 2, 1], bpf_b=[1, 3, 3, 2],
             bpf_f=[1.0, 0.25, 0.75, 1.0], root=[3, 4], profile=np.eye(5, dtype=np.float32)[np.arange(12) % 4],
             n_seqs=1.0, seq_weight=np.linspace(0.1, 1.0, 12))
    a = hostlib.MData.from_arrays(f, "acguacguacgu")
    b = golden_like_hairpin()
    flat = hostlib.SeqSet([a, b])
    for kind in (L.SI_STEM, L.SU_STEM, L.SU_STEM_STR):
        for band in (0, 2):
            p = L.make_params(kind, len_band=band)
            ctx = api.Context(p)
            assert relerr(ctx.gram(ctx.upload(flat)), O.gram(oparams(p), flat.desc(), False)) < TOL


def golden_like_hairpin():
    return hostlib.MData.build(["gggaaaaccc"], [(np.array([1, 2, 3]), np.array([10, 9, 8]), np.array([0.9, 0.8, 0.7]))],
                               TH)


def test_pairs_and_cost_model(golden):
    p = L.make_params(L.SU_STEM_STR)
    ctx = api.Context(p)
    ds = ctx.upload(golden["flat"])
    n = len(ds)
    rng = np.random.default_rng(1)
    xi, yi = rng.integers(0, n, 40), rng.integers(0, n, 40)
    d = golden["flat"].desc()
    assert relerr(ctx.pairs(ds, ds, xi, yi), O.pairs(oparams(p), d, d, xi, yi)) < TOL
    cells, flops = ctx.pair_cost(ds, ds, xi, yi)
    for k in range(40):
        c, f = O.pair_cost(oparams(p), d, int(xi[k]), d, int(yi[k]))
        assert cells[k] == c and flops[k] == f
    with pytest.raises(api.StemkError):
        ctx.pairs(ds, ds, [n], [0])                     # index out of range -> STEMK_ERR_ARG, not a crash
    assert ctx.stats()["launches"] > 0


# ------------------------------------------------------------------ (3) full-size properties
def test_config2_full_size_string_kernel_properties():
    """C2: 2 000 random sequences of 100 nt, 2 001 000 pairs."""
    recs = synth.make_config(2)
    md = [hostlib.MData.seq_only(r["rows"]) for r in recs]
    flat = hostlib.SeqSet(md)
    p = L.make_params(L.STR_SUBST)
    ctx = api.Context(p)
    ds = ctx.upload(flat)
    g = ctx.gram(ds)
    assert g.shape == (2000, 2000) and np.isfinite(g).all() and np.array_equal(g, g.T)
    rng = np.random.default_rng(2)
    xi, yi = rng.integers(0, 2000, 400), rng.integers(0, 2000, 400)
    want = O.pairs(oparams(p), flat.desc(), flat.desc(), np.minimum(xi, yi), np.maximum(xi, yi))
    assert relerr(g[xi, yi], want) < TOL
    gn = ctx.gram(ds, normalize=True)
    assert np.all(np.diag(gn) == 1.0)
    d = np.diag(g)
    assert relerr(gn[xi, yi][xi != yi], (g[xi, yi] / np.sqrt(d[xi] * d[yi]))[xi != yi]) < 1e-15 * 4
    # Cauchy-Schwarz: a Gram matrix of a positive-definite kernel has |k(x,y)| <= sqrt(k(x,x) k(y,y))
    assert np.all(np.abs(gn) <= 1.0 + 1e-12)


def test_config3_stem_kernel_properties_and_sampled_parity():
    """C3 records (150-300 nt): 400 of them -> 80 200 pairs; sampled entries against the oracle; composition
    is exactly additive (AddKernel, conv_kernel.h:49-52); behaviour under a permutation of the records."""
    recs = synth.make_config(3, 400)
    md = hostlib.build_many(recs, TH)
    flat = hostlib.SeqSet(md)
    ps, pss, pstr = L.make_params(L.SU_STEM), L.make_params(L.SU_STEM_STR), L.make_params(L.STR_SUBST)
    cs, css, cstr = api.Context(ps), api.Context(pss), api.Context(pstr)
    gs = cs.gram(cs.upload(flat))
    gss = css.gram(css.upload(flat))
    gstr = cstr.gram(cstr.upload(flat))
    assert np.isfinite(gs).all() and np.array_equal(gs, gs.T)
    assert np.array_equal(gss, gs + gstr)
    rng = np.random.default_rng(3)
    xi, yi = rng.integers(0, 400, 120), rng.integers(0, 400, 120)
    lo, hi = np.minimum(xi, yi), np.maximum(xi, yi)
    assert relerr(gs[lo, hi], O.pairs(oparams(ps), flat.desc(), flat.desc(), lo, hi)) < TOL
    # The stem kernel is NOT symmetric in its arguments (a leaf row of the reference's tables is 0, a leaf column
    # is not: stem_kernel.cpp:39-42,62-77), and KernelMatrix evaluates kernel_(x_i, x_j) with i <= j
    # (kernel_matrix.cpp:47-50).  Under a permutation of the records, entries whose argument order is preserved
    # must not change (beyond the last bits: warps pick up rows dynamically, so the final sum's order may
    # differ between launches); entries whose order flips must equal the oracle with the arguments flipped.
    perm = rng.permutation(400)
    gp = cs.gram(cs.upload([md[i] for i in perm]))
    a, b = np.triu_indices(400)
    kept = perm[a] <= perm[b]
    assert relerr(gp[a[kept], b[kept]], gs[perm[a[kept]], perm[b[kept]]]) < 1e-13
    fl = np.nonzero(~kept)[0][:: max(1, (~kept).sum() // 60)]
    want = O.pairs(oparams(ps), flat.desc(), flat.desc(), perm[a[fl]], perm[b[fl]])
    assert relerr(gp[a[fl], b[fl]], want) < TOL


@pytest.mark.skipif(not os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_svm.so")), reason="oracle/_ref not shipped")
def test_config1_libsvm_cross_validation_is_identical():
    """C1: 200 tRNA-like records, normalised SuStemStrKernel Gram; the vendored LIBSVM's 5-fold CV targets on the
    GPU matrix equal those on the oracle's matrix (north_star: identical cross-validation predictions)."""
    recs = synth.make_config(1)
    md = hostlib.build_many(recs, TH)
    flat = hostlib.SeqSet(md)
    p = L.make_params(L.SU_STEM_STR)
    ctx = api.Context(p)
    got = ctx.gram(ctx.upload(flat), normalize=True)
    want = O.gram(oparams(p), flat.desc(), True)
    assert relerr(got, want) < TOL
    y = np.array([r["label"] for r in recs], dtype=np.float64)
    assert np.array_equal(R.svm_cv(got, y, 1.0, 5, 1), R.svm_cv(want, y, 1.0, 5, 1))
    z = np.load(os.path.join(GOLDEN, "golden_svm.npz"))
    assert relerr(got[:40, :40], z["gram"]) < TOL


def test_long_records_take_the_general_kernel_next_to_fast_ones():
    """500-nt records: pairs up to ~490 nt apart push the fast path's gap powers g^(+-len/2) out of its safe range
    and the DAGs are big (~1000 nodes), so these records must run on the general kernel while the 150-300 nt
    records of the same set keep the fast one -- both inside one stemk_gram call."""
    recs = [synth.ncrna_like(424242, i, lmin=500, lmax=500) for i in range(3)] + synth.make_config(3, 4, offset=8800)
    md = [hostlib.MData.from_record(r, TH) for r in recs]
    assert max(m.sizes()["n_nodes"] for m in md[:3]) > 700
    flat = hostlib.SeqSet(md)
    for kind in (L.SU_STEM, L.SU_STEM_STR):
        p = L.make_params(kind)
        ctx = api.Context(p)
        got = ctx.gram(ctx.upload(flat))
        assert relerr(got, O.gram(oparams(p), flat.desc(), False)) < TOL


def test_record_too_large_for_shared_memory_takes_the_unstaged_kernel():
    """A 1 500-nt record (2 500 nodes, 17 000 edges) exceeds what the general kernel can stage in shared memory; until
    round 2 the call failed with STEMK_ERR_NOMEM, now the pair runs on the unstaged kernel and must equal the oracle."""
    recs = [synth.ncrna_like(77, 0, lmin=1500, lmax=1500)]
    md = [hostlib.MData.from_record(r, 0.001) for r in recs]      # low threshold: every background pair is a node
    assert md[0].sizes()["n_nodes"] > 2000
    flat = hostlib.SeqSet(md)
    p = L.make_params(L.SU_STEM)
    ctx = api.Context(p)
    got = ctx.gram(ctx.upload(flat))
    assert relerr(got, O.gram(oparams(p), flat.desc(), False)) < TOL


def test_rectangular_matrix_with_sv_subset_at_moderate_size():
    """KernelMatrix::calculate(test, train) / the one-row variant with an sv_index (kernel_matrix.cpp:635-754):
    60 test x 150 train C3 records, 40 support-vector columns, normalised; sampled entries against the oracle,
    columns outside sv_index untouched, NaN pattern of the normalisation as in App::predict."""
    train = hostlib.build_many(synth.make_config(3, 150, offset=3000), TH)
    test = hostlib.build_many(synth.make_config(3, 60, offset=5000), TH)
    ftrain, ftest = hostlib.SeqSet(train), hostlib.SeqSet(test)
    p = L.make_params(L.SU_STEM_STR)
    ctx = api.Context(p)
    dtrain, dtest = ctx.upload(ftrain), ctx.upload(ftest)
    rng = np.random.default_rng(9)
    sv = np.sort(rng.choice(150, 40, replace=False)).astype(np.uint32)
    m, selfv = ctx.cross(dtest, dtrain, sv_index=sv, init=-7.0)
    others = np.setdiff1d(np.arange(150), sv)
    assert np.all(m[:, others] == -7.0)
    ti, ci = rng.integers(0, 60, 150), sv[rng.integers(0, 40, 150)]
    want = O.pairs(oparams(p), ftrain.desc(), ftest.desc(), ci, ti)       # train record is the first argument
    assert relerr(m[ti, ci], want) < TOL
    assert relerr(selfv[:10], O.pairs(oparams(p), ftest.desc(), ftest.desc(), np.arange(10), np.arange(10))) < TOL
    full, _ = ctx.cross(dtest, dtrain)
    assert relerr(full[:, sv], m[:, sv]) == 0.0
    mn, _ = ctx.cross(dtest, dtrain, normalize=True)
    d = ctx.diag(dtrain)
    assert relerr(mn, full / np.sqrt(np.outer(selfv, d))) < 1e-15 * 8


@pytest.mark.parametrize("case", ["all_weighted", "none_weighted", "mixed", "gap_columns", "wide_weighted"])
def test_string_kernel_launch_modes(case):
    """The string kernel is specialised per launch (plain / weighted / general): every mode against the oracle."""
    recs = synth.make_config(1, 6, offset=7000)
    if case == "all_weighted":
        md = [hostlib.MData.from_record(r, TH) for r in recs]
    elif case == "none_weighted":
        md = [hostlib.MData.seq_only(r["rows"]) for r in recs]
    elif case == "mixed":
        md = [hostlib.MData.from_record(r, TH) if i % 2 else hostlib.MData.seq_only(r["rows"]) for i, r in enumerate(recs)]
    elif case == "gap_columns":   # a one-row record with '-' columns: one-hot-or-gap columns, score 1 on the gaps
        md = [hostlib.MData.seq_only([r["rows"][0][:20] + "--" + r["rows"][0][20:40] + "-" + r["rows"][0][40:]]) for r in recs]
    else:
        md = hostlib.build_many(synth.make_config(3, 5, offset=7100), TH)     # 150-300 columns, weighted
    flat = hostlib.SeqSet(md)
    for kind in (L.STR_SUBST, L.STR_SIMPLE):
        p = L.make_params(kind)
        ctx = api.Context(p)
        assert relerr(ctx.gram(ctx.upload(flat)), O.gram(oparams(p), flat.desc(), False)) < TOL


def test_sharded_driver_single_rank_equals_gram():
    import torch
    from stem_kernel_b200 import sharded
    recs = synth.make_config(3, 48)
    md = hostlib.build_many(recs, TH)
    ctx = api.Context(L.make_params(L.SU_STEM))
    ds = ctx.upload(md)
    dev = torch.device("cuda", 0)
    be = sharded.GpuBackend(ctx, ds, dev)
    sg = sharded.ShardedGram(sharded.record_keys(ds), 0, 1, dev, be.compute, be.assemble)
    with torch.cuda.stream(be.stream):
        m = sg.run(normalize=True)
    be.stream.synchronize()
    assert np.array_equal(m.cpu().numpy(), ctx.gram(ds, normalize=True), equal_nan=True)


def test_stream_predict_writes_the_reference_text():
    """Streaming predict + writer (App::predict / Output, common/framework.h:167-306,
    framework.cpp:141-234): test rows computed batch by batch, normalised like framework.h:279-283, written by the
    native threaded writer.  The files must be byte-identical to formatting the streamed rows with the Python
    mirror of KernelMatrix::print / Output::kernel_output, and the rows must agree with the one-shot normalised
    cross matrix to 1e-12 whatever the batch size (the string kernel's tile shape follows the batch, so the last
    bits of a row may differ between batch sizes)."""
    import io
    train = hostlib.build_many(synth.make_config(1, 40, offset=100), TH)
    test_recs = synth.make_config(1, 23, offset=900)
    test = hostlib.build_many(test_recs, TH)
    ctx = api.Context(L.make_params(L.SU_STEM_STR))
    dtrain = ctx.upload(hostlib.SeqSet(train))
    labels = ["+1" if i % 3 else "-1" for i in range(len(test))]
    want_m, want_self = ctx.cross(ctx.upload(hostlib.SeqSet(test)), dtrain, normalize=True)
    want = api.format_matrix(want_m, labels).encode()
    want_norm = "".join(api._g6(v) + "\n" for v in want_self).encode()
    for bs in (1, 5, 23):
        def batches():
            for b in range(0, len(test), bs):
                yield labels[b:b + bs], ctx.upload(hostlib.SeqSet(test[b:b + bs]))
        out, nout = io.BytesIO(), io.BytesIO()
        seen, got_rows, got_self = 0, [], []
        for lab, rows, selfv in api.stream_predict(ctx, dtrain, batches(), out, norm_out=nout, normalize=True):
            assert rows.shape == (len(lab), len(train))
            np.testing.assert_allclose(rows, want_m[seen:seen + len(lab)], rtol=1e-12, atol=0)
            np.testing.assert_allclose(selfv, want_self[seen:seen + len(lab)], rtol=1e-12, atol=0)
            got_rows.append(rows); got_self.append(selfv)
            seen += len(lab)
        assert seen == len(test)
        assert out.getvalue() == api.format_matrix(np.concatenate(got_rows), labels).encode()
        assert nout.getvalue() == "".join(api._g6(v) + "\n" for v in np.concatenate(got_self)).encode()
        if bs == len(test):
            assert out.getvalue() == want and nout.getvalue() == want_norm


@pytest.mark.skipif(not os.path.exists(os.path.join(R.REF_DIR, "libstemk_ref_svm.so")), reason="oracle/_ref not shipped")
def test_sharded_cross_feeds_identical_svm_predictions():
    """BASELINE config 5 at test size: rectangular test x train matrix through the multi-GPU driver's rectangular path
    (ShardedCross with world 1: pair lists, self terms, train diagonals, one buffer, scatter, normalisation) against
    the oracle and against stemk_cross; the vendored LIBSVM trained on the normalised train Gram predicts the same
    labels from the GPU rows as from the oracle's rows (north_star: identical downstream predictions)."""
    import torch
    from stem_kernel_b200 import sharded
    rtrain, rtest = synth.make_config(1, 60, offset=300), synth.make_config(1, 25, offset=700)
    train, test = hostlib.build_many(rtrain, TH), hostlib.build_many(rtest, TH)
    ftrain, ftest = hostlib.SeqSet(train), hostlib.SeqSet(test)
    p = L.make_params(L.SU_STEM_STR)
    ctx = api.Context(p)
    dtrain, dtest = ctx.upload(ftrain), ctx.upload(ftest)
    dev = torch.device("cuda", 0)
    be = sharded.GpuCrossBackend(ctx, dtrain, dtest, dev)
    sc = sharded.ShardedCross(sharded.record_keys(dtest, string=True), sharded.record_keys(dtrain, string=True), 0, 1, dev,
                              be.compute)
    assert sc.n_pairs == 25 * 60 + 25 + 60
    with torch.cuda.stream(be.stream):
        m, selfv = sc.run(normalize=True)
    be.stream.synchronize()
    got, got_self = m.cpu().numpy(), selfv.cpu().numpy()
    want, want_self = O.cross(oparams(p), ftest.desc(), ftrain.desc(), normalize=True)
    assert relerr(got, want) < TOL and relerr(got_self, want_self) < TOL
    one, one_self = ctx.cross(dtest, dtrain, normalize=True)
    np.testing.assert_allclose(got, one, rtol=1e-12, atol=0)
    np.testing.assert_allclose(got_self, one_self, rtol=1e-12, atol=0)
    ktrain = ctx.gram(dtrain, normalize=True)
    y = np.array([r["label"] for r in rtrain], dtype=np.float64)
    pred_gpu = R.svm_train_predict(ktrain, y, got)
    pred_ref = R.svm_train_predict(O.gram(oparams(p), ftrain.desc(), True), y, want)
    assert np.array_equal(pred_gpu, pred_ref) and set(np.unique(pred_gpu)) <= {-1.0, 1.0}
