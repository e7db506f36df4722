"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d): random RNA with planted hairpins and a
deterministic stand-in for the McCaskill base-pair probabilities (no ViennaRNA in this image).

Everything is a function of (seed, sequence index) through numpy's PCG64, so the reference arm, the
oracle and the CUDA path are fed identical sequences and identical thresholded base-pair lists.

A sequence record is a dict:
    rows   list[str]   aligned rows (one row for a plain FASTA record), alphabet acgu (+ '-' / IUPAC)
    bp     list[(bi, bj, bp)]  per row, sparse probabilities over the UNGAPPED row, 1-based, bi < bj
    label  int
"""
import numpy as np

BASES = np.array(list("acgu"))
_COMP = {("a", "u"), ("u", "a"), ("g", "c"), ("c", "g"), ("g", "u"), ("u", "g")}
_WC = {"a": "u", "u": "a", "g": "c", "c": "g"}

# complementarity table on codes a=0,c=1,g=2,u=3
_CM = np.zeros((4, 4), dtype=bool)
for _a, _b in _COMP:
    _CM["acgu".index(_a), "acgu".index(_b)] = True


def _rng(seed, idx, stream=0):
    return np.random.Generator(np.random.PCG64([int(seed), int(idx), int(stream)]))


def _plant(codes, stems, rng):
    """Overwrite the 3' arm of each (i, j, n) stem (0-based outer pair i<j, n stacked pairs) with the
    Watson-Crick complement of its 5' arm; return the planted pairs (0-based)."""
    pairs = []
    comp = np.array([3, 2, 1, 0])
    for (i, j, n) in stems:
        for k in range(n):
            codes[j - k] = comp[codes[i + k]]
            pairs.append((i + k, j - k))
    return pairs


def stand_in_bp(codes, planted, rng, bg_per_nt=1.6, p_planted=(0.6, 0.95), p_bg=(0.01, 0.3), min_span=4):
    """Deterministic stand-in for pf_fold's pair probabilities: planted pairs get
    p in [0.6,0.95], a random subset of the other complementary pairs (j-i >= min_span) gets
    p in [0.01,0.3]; then every position's total pairing probability is scaled down to <= 1.
    Returns 1-based sparse (bi, bj, bp) sorted by (bj, bi)."""
    L = len(codes)
    P = {}
    for (i, j) in planted:
        P[(i, j)] = rng.uniform(*p_planted)
    ii, jj = np.triu_indices(L, k=min_span)
    ok = _CM[codes[ii], codes[jj]]
    ii, jj = ii[ok], jj[ok]
    n_bg = min(len(ii), int(round(bg_per_nt * L)))
    if n_bg > 0:
        sel = rng.choice(len(ii), size=n_bg, replace=False)
        pb = rng.uniform(*p_bg, size=n_bg)
        for s, p in zip(sel, pb):
            key = (int(ii[s]), int(jj[s]))
            if key not in P:
                P[key] = float(p)
    if not P:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros(0)
    keys = np.array(sorted(P.keys(), key=lambda k: (k[1], k[0])), dtype=np.int64)
    vals = np.array([P[(a, b)] for a, b in keys])
    # scale so that sum over partners <= 1 at every position (one pass suffices: scaling only shrinks)
    for pos in range(L):
        m = (keys[:, 0] == pos) | (keys[:, 1] == pos)
        s = vals[m].sum()
        if s > 1.0:
            vals[m] *= 0.999 / s
    return keys[:, 0] + 1, keys[:, 1] + 1, vals


def _record(codes, planted, rng, label, **kw):
    seq = "".join(BASES[codes])
    return dict(rows=[seq], bp=[stand_in_bp(codes, planted, rng, **kw)], label=label)


def trna_like(seed, idx):
    """C1 record: L ~ U{70..80}, cloverleaf of 4 planted stems (acceptor 7, D 4, anticodon 5, T 5)."""
    rng = _rng(seed, idx)
    L = int(rng.integers(70, 81))
    codes = rng.integers(0, 4, size=L)
    # layout: acceptor closes everything; three hairpin arms inside
    d_loop, a_loop, t_loop = int(rng.integers(7, 9)), 7, 7
    p = 7 + 2                                   # after the acceptor 5' arm + 2 nt
    d = (p, p + 2 * 4 + d_loop - 1, 4)
    p = d[1] + 2
    a = (p, p + 2 * 5 + a_loop - 1, 5)
    p = a[1] + 1 + (L - 76 if L > 76 else 0) + 3  # variable region
    t = (p, p + 2 * 5 + t_loop - 1, 5)
    acc_j = L - 2
    stems = [(0, acc_j, 7)]
    for s in (d, a, t):
        if s[1] < acc_j - 7:
            stems.append(s)
    planted = _plant(codes, stems, rng)
    return _record(codes, planted, rng, +1 if idx % 2 == 0 else -1)


def random_seq(seed, idx, L=100):
    """C2 record: iid uniform acgu, no structure (sequence-only; bp list empty)."""
    rng = _rng(seed, idx)
    codes = rng.integers(0, 4, size=L)
    z = np.zeros(0, dtype=np.int64)
    return dict(rows=["".join(BASES[codes])], bp=[(z, z, np.zeros(0))], label=+1 if idx % 2 == 0 else -1)


def ncrna_like(seed, idx, lmin=150, lmax=300):
    """C3/C4/C5 record: L ~ U{lmin..lmax}; 2-5 planted hairpins (stems 4-10 bp, loops 4-8 nt), the first
    k of them optionally enclosed by a multiloop-closing stem."""
    rng = _rng(seed, idx)
    L = int(rng.integers(lmin, lmax + 1))
    codes = rng.integers(0, 4, size=L)
    n_hp = int(rng.integers(2, 6))
    stems = []
    # place hairpins left to right in disjoint windows
    win = (L - 20) // n_hp
    for h in range(n_hp):
        n = int(rng.integers(4, 11))
        loop = int(rng.integers(4, 9))
        span = 2 * n + loop
        if span + 2 > win:
            n = max(4, (win - loop - 2) // 2)
            span = 2 * n + loop
        lo = 10 + h * win
        start = lo + int(rng.integers(0, max(1, win - span)))
        stems.append((start, start + span - 1, n))
    if rng.random() < 0.5 and n_hp >= 2:
        n = int(rng.integers(4, 8))
        i = max(0, stems[0][0] - n - int(rng.integers(1, 3)))
        j = min(L - 1, stems[-1][1] + n + int(rng.integers(1, 3)))
        if i + n <= stems[0][0] and j - n >= stems[-1][1]:
            stems.insert(0, (i, j, n))
    planted = _plant(codes, stems, rng)
    return _record(codes, planted, rng, +1 if idx % 2 == 0 else -1)


def alignment_like(seed, idx, n_rows=3, L=60, gap_rate=0.08, iupac_rate=0.03):
    """Edge-case record: a gapped alignment of n_rows rows sharing one planted hairpin, with a few IUPAC
    codes (exercises multi-entry bp_freq lists and the gap terms of node_score, score_table.cpp:185-198)."""
    rng = _rng(seed, idx, 7)
    base = rng.integers(0, 4, size=L)
    n = int(rng.integers(4, 8))
    loop = int(rng.integers(4, 8))
    start = int(rng.integers(2, L - 2 * n - loop - 2))
    stem = (start, start + 2 * n + loop - 1, n)
    rows, bps = [], []
    iupac = "rymkswbdhvn"
    for r in range(n_rows):
        codes = base.copy()
        mut = rng.random(L) < 0.1
        codes[mut] = rng.integers(0, 4, size=int(mut.sum()))
        planted = _plant(codes, [stem], rng)
        chars = list(BASES[codes])
        gaps = rng.random(L) < gap_rate
        if r == 0:
            gaps[:] = False
        for k in np.nonzero(rng.random(L) < iupac_rate)[0]:
            chars[k] = iupac[int(rng.integers(0, len(iupac)))]
        row = "".join("-" if g else c for c, g in zip(chars, gaps))
        # probabilities live on the ungapped row
        keep = np.nonzero(~gaps)[0]
        remap = -np.ones(L, dtype=np.int64)
        remap[keep] = np.arange(len(keep))
        ung_codes = codes[keep]
        ung_planted = [(int(remap[a]), int(remap[b])) for a, b in planted if remap[a] >= 0 and remap[b] >= 0
                       and remap[b] - remap[a] >= 4]
        rows.append(row)
        bps.append(stand_in_bp(ung_codes, ung_planted, rng, bg_per_nt=1.0))
    return dict(rows=rows, bp=bps, label=+1 if idx % 2 == 0 else -1)


CONFIG_SEED = {1: 20260001, 2: 20260002, 3: 20260003, 4: 20260004, 5: 20260005}


def make_config(cfg, n=None, offset=0):
    """Records of BASELINE config `cfg` (1..5). n overrides the named size; offset shifts the index
    stream (C5's test set uses a disjoint index range of the same generator)."""
    seed = CONFIG_SEED[cfg]
    if cfg == 1:
        n = 200 if n is None else n
        return [trna_like(seed, offset + i) for i in range(n)]
    if cfg == 2:
        n = 2000 if n is None else n
        return [random_seq(seed, offset + i, 100) for i in range(n)]
    if cfg in (3, 4, 5):
        n = {3: 2000, 4: 10000, 5: 10000}[cfg] if n is None else n
        return [ncrna_like(seed, offset + i) for i in range(n)]
    raise ValueError(cfg)
