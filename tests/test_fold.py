"""Base-pair-probability front end: stemk_fold_bpp against the oracle restatement
(oracle/stemk_fold_oracle.c), and the oracle against an exhaustive enumeration of every secondary structure of short
sequences under the same loop model, written here independently of both.

PARITY UNPINNED: the reference takes these probabilities from ViennaRNA (common/bpmatrix.cpp:141-177), which is
absent from this image and from /root/reference; nothing in the reference pins a value at that boundary.  What these
tests pin is (1) that the McCaskill recursions of the oracle sum exactly the Boltzmann ensemble of the stated model,
(2) that the CUDA path equals the oracle to 1e-9."""
import math

import numpy as np
import pytest

from conftest import need_gpu, relerr
from oracle import oraclebind as O

KT37 = (37.0 + 273.15) * 1.98717e-3
CODE = {"a": 1, "c": 2, "g": 3, "u": 4, "t": 4}
PAIR = {(2, 3): 1, (3, 2): 2, (3, 4): 3, (4, 3): 4, (1, 4): 5, (4, 1): 6}
RTYPE = [0, 2, 1, 4, 3, 6, 5]


def codes(seq):
    return [0] + [CODE.get(c.lower(), 0) for c in seq] + [0]


def structures(S, n, no_gu=False):
    """Every non-crossing set of allowed pairs on 1..n (at least 3 unpaired bases inside a pair)."""
    def ptype(i, j):
        t = PAIR.get((S[i], S[j]), 0)
        return 0 if (no_gu and t in (3, 4)) else t

    def rec(i, j):
        if j - i < 4:
            yield ()
            return
        for s in rec(i + 1, j):
            yield s
        for k in range(i + 4, j + 1):
            if ptype(i, k):
                for a in rec(i + 1, k - 1):
                    for b in rec(k + 1, j):
                        yield ((i, k),) + a + b
    return rec(1, n), ptype


def structure_energy(m, S, n, pairs, ptype):
    """Loop decomposition of one structure under the model of include/stemk.h (stemk_fold_model)."""
    pairs = sorted(pairs)
    children = {p: [] for p in pairs}
    top = []
    stack = []
    for p in pairs:
        while stack and stack[-1][1] < p[0]:
            stack.pop()
        (children[stack[-1]] if stack else top).append(p)
        stack.append(p)
    e = 0.0
    for (i, j) in top:
        t = ptype(i, j)
        e += (m.terminal_au if t > 2 else 0.0) + (m.dangle5[t][S[i - 1]] if i > 1 else 0.0) + (m.dangle3[t][S[j + 1]] if j < n else 0.0)
    for (i, j), ch in children.items():
        t = ptype(i, j)
        if not ch:
            u = j - i - 1
            e += m.hairpin[u] if u <= 30 else m.hairpin[30] + m.lxc * math.log(u / 30.0)
            e += (m.terminal_au if t > 2 else 0.0) if u == 3 else m.mismatch_h[t][S[i + 1]][S[j - 1]]
        elif len(ch) == 1:
            (k, l) = ch[0]
            t2 = RTYPE[ptype(k, l)]
            u1, u2 = k - i - 1, j - l - 1
            assert u1 + u2 <= 30
            if u1 == 0 and u2 == 0:
                e += m.stack[t][t2]
            elif u1 == 0 or u2 == 0:
                e += m.bulge[u1 + u2]
                e += m.stack[t][t2] if u1 + u2 == 1 else (m.terminal_au if t > 2 else 0.0) + (m.terminal_au if t2 > 2 else 0.0)
            else:
                e += m.interior[u1 + u2] + min(m.max_ninio, abs(u1 - u2) * m.ninio)
                e += m.mismatch_i[t][S[i + 1]][S[j - 1]] + m.mismatch_i[t2][S[l + 1]][S[k - 1]]
        else:
            tt = RTYPE[t]
            e += m.ml_closing + m.ml_intern[tt] + m.dangle3[tt][S[i + 1]] + m.dangle5[tt][S[j - 1]]
            unp = j - i - 1
            for (k, l) in ch:
                tc = ptype(k, l)
                e += m.ml_intern[tc] + m.dangle5[tc][S[k - 1]] + m.dangle3[tc][S[l + 1]]
                unp -= l - k + 1
            e += m.ml_base * unp
    return e


def enumerate_bpp(m, seq):
    n = len(seq)
    S = codes(seq)
    kT = (m.temperature + 273.15) * 1.98717e-3
    gen, ptype = structures(S, n, bool(m.no_gu))
    Z = 0.0
    P = np.zeros((n + 1, n + 1))
    count = 0
    for st in gen:
        w = math.exp(-structure_energy(m, S, n, st, ptype) / kT)
        Z += w
        for (i, j) in st:
            P[i, j] += w
        count += 1
    return P / Z, -kT * math.log(Z), count


def random_model(seed, temperature=25.0):
    """A loop model with every table entry drawn at random (seeded): a swapped index or a transposed table in any of
    the three implementations (enumeration here, oracle, CUDA) shows up as a mismatch."""
    rng = np.random.default_rng(seed)
    m = O.fold_model_default()
    m.temperature = temperature
    for t in range(1, 7):
        for u in range(1, 7):
            m.stack[t][u] = float(rng.uniform(-3.5, 0.5))
        m.ml_intern[t] = float(rng.uniform(0.0, 1.0))
        for a in range(5):
            m.dangle5[t][a] = float(rng.uniform(-0.8, 0.0))
            m.dangle3[t][a] = float(rng.uniform(-0.8, 0.0))
            for b in range(5):
                m.mismatch_h[t][a][b] = float(rng.uniform(-1.5, 0.0))
                m.mismatch_i[t][a][b] = float(rng.uniform(-1.0, 0.7))
    for u in range(31):
        m.hairpin[u] = float(rng.uniform(3.0, 7.0))
        m.bulge[u] = float(rng.uniform(2.0, 6.0))
        m.interior[u] = float(rng.uniform(1.0, 5.0))
    m.ninio, m.max_ninio = float(rng.uniform(0.2, 0.8)), float(rng.uniform(1.0, 3.0))
    m.terminal_au = float(rng.uniform(0.0, 1.0))
    m.ml_closing, m.ml_base = float(rng.uniform(1.0, 4.0)), float(rng.uniform(0.0, 0.3))
    return m


SHORT = ["gggaaaccc", "gcgcuuuugcgc", "ggaaacgaaacgcc", "gggaaauccgaaaggaaaccc"[:18], "gacuuagguuaccgagucu"[:17],
         "gugucgaaagacgaaaguc"[:18], "ggnaaaccc", "aaaaaaaa", "gcgaugcuuagc", "ggcgaaagccgaaaggc",
         # 7 481 / 11 029 / 13 043 / 42 860 structures: interior loops, bulges and multiloops with several branch layouts
         "gggcgcaagcgcgcaagcgccc", "gcgguuagcgcaaugcgcuagc", "ggugcgaaagcaugcgaaagcacc", "gcaugcuuuugcagcuuaagcugc"]


@pytest.mark.parametrize("seq", SHORT)
def test_oracle_equals_exhaustive_enumeration(seq):
    m = O.fold_model_default()
    m.ml_base = 0.15            # exercise the per-base multiloop term as well
    want, ens, count = enumerate_bpp(m, seq)
    got, gens, unp = O.fold_bpp(m, seq)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-13), (seq, count)
    assert abs(gens - ens) < 1e-9
    assert np.allclose(unp, np.maximum(0.0, 1.0 - want[1:, 1:].sum(0) - want[1:, 1:].sum(1)), atol=1e-12)


@pytest.mark.parametrize("seq", ["gggcgcaagcgcgcaagcgccc", "gcgguuagcgcaaugcgcuagc", "gacuuagguuaccgagu", "ggugcgaaagcaugcgaaagcacc"])
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_equals_exhaustive_enumeration_random_model(seq, seed):
    m = random_model(seed)
    want, ens, _ = enumerate_bpp(m, seq)
    got, gens, _ = O.fold_bpp(m, seq)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-13) and abs(gens - ens) < 1e-9


def test_enumeration_covers_multiloops():
    """The short set must contain structures with a multiloop, otherwise the check above says nothing about qm/qm1."""
    m = O.fold_model_default()
    seq = "ggcgaaagccgaaaggc"
    S = codes(seq)
    gen, ptype = structures(S, len(seq))
    assert any(sum(1 for (k, l) in st if st[0][0] < k and l < st[0][1]) >= 2 and (1, len(seq)) == st[0] for st in gen if st)


def test_oracle_scale_invariance_and_options():
    rng = np.random.default_rng(5)
    seq = "".join("acgu"[c] for c in rng.integers(0, 4, 120))
    m = O.fold_model_default()
    base, ens, _ = O.fold_bpp(m, seq)
    for s in (1.0, 1.2, 1.6):
        m.pf_scale = s
        got, e2, _ = O.fold_bpp(m, seq)
        assert np.allclose(got, base, rtol=1e-9, atol=1e-14) and abs(e2 - ens) < 1e-8
    m.pf_scale = -1.0
    m.no_gu = 1
    got, _, unp = O.fold_bpp(m, seq)
    S = codes(seq)
    for i, j in zip(*np.nonzero(got)):
        assert PAIR[(S[i], S[j])] not in (3, 4)
    assert np.all(got.sum(0) + got.sum(1) <= 1.0 + 1e-12) and np.all(unp >= 0.0)
    assert O.fold_bpp(m, "")[0].shape == (1, 1)


# ------------------------------------------------------------------------------------------------ CUDA path
def _random_seqs(seed, lengths):
    rng = np.random.default_rng(seed)
    return ["".join("acgu"[c] for c in rng.integers(0, 4, n)) for n in lengths]


@pytest.mark.gpu
def test_device_fold_equals_oracle():
    need_gpu()
    from stem_kernel_b200 import fold
    seqs = SHORT + _random_seqs(11, [0, 1, 4, 5, 33, 64, 97, 150, 211, 300]) + ["GGGAAATCCCNNNGGGTTTCCC", "acgu" * 40]
    m = fold.default_model()
    m.ml_base = 0.1
    with fold.Folder() as f:
        res = f.bpp(seqs, m, cutoff=0.0, dense=True)
        for k, s in enumerate(seqs):
            want, ens, unp = O.fold_bpp(m, s)
            assert relerr(res.dense[k], want) < 1e-9, (k, len(s))
            assert abs(res.ensemble[k] - ens) <= 1e-9 * max(1.0, abs(ens))
            assert np.allclose(res.unpaired[k], unp, atol=1e-11)
            i, j, p = res.pairs[k]
            wi, wj = np.nonzero(want)
            assert np.array_equal(i, wi) and np.array_equal(j, wj) and relerr(p, want[wi, wj]) < 1e-9
        # thresholded lists: exactly the pairs at or above the cut-off, in ascending (i, j)
        res = f.bpp(seqs, m, cutoff=0.01)
        for k, s in enumerate(seqs):
            want = O.fold_bpp(m, s)[0]
            i, j, p = res.pairs[k]
            sure = want >= 0.01 * (1 + 1e-9)
            maybe = want >= 0.01 * (1 - 1e-9)
            got = np.zeros_like(want, dtype=bool)
            got[i, j] = True
            assert np.all(got[sure]) and not np.any(got & ~maybe)
            assert np.all(np.diff(i * (len(s) + 1) + j) > 0)
        # a model with random tables at another temperature
        rm = random_model(7, temperature=30.0)
        rm2 = fold.FoldModel.from_buffer_copy(rm)
        res = f.bpp(seqs[-6:], rm2, cutoff=0.0, dense=True)
        for k, s in enumerate(seqs[-6:]):
            want, ens, _ = O.fold_bpp(rm, s)
            assert relerr(res.dense[k], want) < 1e-9 and abs(res.ensemble[k] - ens) <= 1e-9 * max(1.0, abs(ens))
        # scaling and the no-GU switch
        m.pf_scale = 1.1
        m.no_gu = 1
        res = f.bpp(seqs[-4:], m, cutoff=0.0, dense=True)
        for k, s in enumerate(seqs[-4:]):
            assert relerr(res.dense[k], O.fold_bpp(m, s)[0]) < 1e-9


@pytest.mark.gpu
def test_device_fold_long_and_many():
    """Longer than any shared-memory shortcut (700 nt) and a batch with more sequences than CTAs."""
    need_gpu()
    from stem_kernel_b200 import fold
    m = fold.default_model()
    seqs = _random_seqs(3, [700]) + _random_seqs(4, [60 + (k % 50) for k in range(400)])
    with fold.Folder() as f:
        res = f.bpp(seqs, m, cutoff=1e-4)
        for k in (0, 1, 57, 399):
            want = O.fold_bpp(m, seqs[k])[0]
            i, j, p = res.pairs[k]
            assert relerr(p, want[i, j]) < 1e-9
            assert (want >= 1e-4 * (1 + 1e-9)).sum() <= len(p) <= (want >= 1e-4 * (1 - 1e-9)).sum()


@pytest.mark.gpu
def test_fold_feeds_the_front_end_and_the_kernel():
    """The pipeline of the reference's Data constructor with the device fold in the place of ViennaRNA: probabilities ->
    Profiler + DAGBuilder (host/frontend.cpp) -> stem kernel Gram matrix, equal to the same pipeline run from the oracle's
    probabilities."""
    need_gpu()
    from stem_kernel_b200 import _lib as L
    from stem_kernel_b200 import api, fold, hostlib
    seqs = _random_seqs(21, [80, 95, 120, 77, 101])
    m = fold.default_model()
    with fold.Folder() as f:
        res = f.bpp(seqs, m, cutoff=0.0)
    md_dev, md_cpu = fold.build_mdata(seqs, 0.01, m), []
    for k, s in enumerate(seqs):
        i, j, p = res.pairs[k]
        assert hostlib.MData.from_record(dict(rows=[s], bp=[(i, j, p)], label=1), 0.01).sizes() == md_dev[k].sizes()
        want = O.fold_bpp(m, s)[0]
        wi, wj = np.nonzero(want)
        md_cpu.append(hostlib.MData.from_record(dict(rows=[s], bp=[(wi, wj, want[wi, wj])], label=1), 0.01))
    ctx = api.Context(L.make_params(L.SU_STEM_STR))
    a = ctx.gram(ctx.upload(md_dev))
    b = ctx.gram(ctx.upload(md_cpu))
    assert [x.sizes() for x in md_dev] == [x.sizes() for x in md_cpu]
    assert relerr(a, b) < 1e-7      # probabilities within 1e-9 pass through float weights and products of up to ~100 of them
    ctx.close()
