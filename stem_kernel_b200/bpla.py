"""BPLA / local-alignment kernels: Python mirror of bpla_kernel/ over the C ABI.

    BPLAKernel(score_table, noBP, SW, gap, ext, alpha, beta)      bpla_kernel/bpla_kernel.h:13-46
    Data = {seq: ProfileSequence, p_left, p_right, p_unpair}      bpla_kernel/data.h:30-53

`BplaSet` is a flattened vector<Data>; `pairing_profiles` restates data.cpp:19-46 (the part of the Data constructor
after the external ViennaRNA call) on a thresholded base-pair list.  The kernel values come from the CUDA library
only (stemk_bpla_pairs); there is no CPU path here."""
import ctypes as C

import numpy as np

from . import _lib as L
from . import hostlib

DEFAULT_SCORE = np.array([[5.846613, -1.860000, -1.460000, -1.390000],      # bpla_kernel/main.cpp:20-26 (float literals)
                          [-1.860000, 4.786613, -2.480000, -1.050000],
                          [-1.460000, -2.480000, 4.656613, -1.740000],
                          [-1.390000, -1.050000, -1.740000, 5.276613]], dtype=np.float32).astype(np.float64)


class BplaParams(C.Structure):
    """stemk_bpla_params.  Defaults: bpla_kernel/main.cpp:69-73 (gap and ext are parsed as float there)."""
    _fields_ = [("no_bp", C.c_int32), ("sw", C.c_int32), ("gap", C.c_double), ("ext", C.c_double), ("alpha", C.c_double),
                ("beta", C.c_double), ("score", C.c_double * 16)]


def make_params(no_bp=False, sw=False, gap=-8.0, ext=-0.75, alpha=4.5, beta=0.11, score=None):
    f = lambda v: float(np.float32(v))
    p = BplaParams(int(no_bp), int(sw), f(gap), f(ext), f(alpha), f(beta))
    t = DEFAULT_SCORE if score is None else np.asarray(score, dtype=np.float64)
    for k in range(16):
        p.score[k] = float(t.reshape(-1)[k])
    return p


class BplaSetC(C.Structure):
    """stemk_bpla_set."""
    _fields_ = [("n_seqs", C.c_uint32), ("col_off", C.c_void_p), ("profile", C.c_void_p), ("p_left", C.c_void_p),
                ("p_right", C.c_void_p), ("p_unpair", C.c_void_p)]


def pairing_profiles(length, bp):
    """p_left, p_right, p_unpair of one sequence from its base-pair list (bi, bj, p), 1-based, bi < bj:
    data.cpp:19-46 (float accumulators fed with double probabilities, column by column in ascending partner order,
    clamp at 0, square roots in float)."""
    bi, bj, pp = (np.asarray(a) for a in bp)
    pl, pr = np.zeros(length, dtype=np.float32), np.zeros(length, dtype=np.float32)
    order = np.lexsort((bi, bj))                      # p_right[i]: partners j < i ascending
    for k in order:
        i = int(bj[k]) - 1
        pr[i] = np.float32(np.float64(pr[i]) + np.float64(pp[k]))
    order = np.lexsort((bj, bi))                      # p_left[i]: partners j > i ascending
    for k in order:
        i = int(bi[k]) - 1
        pl[i] = np.float32(np.float64(pl[i]) + np.float64(pp[k]))
    pu = (np.float64(1.0) - (pl + pr).astype(np.float64)).astype(np.float32)     # 1.0-(p_l+p_r): float sum, double subtraction
    pu[pu < 0] = 0
    return np.sqrt(pl), np.sqrt(pr), np.sqrt(pu)


class BplaSet:
    """Flattened vector<Data>.  records: list of dicts {"rows": [aligned strings], "p_left", "p_right", "p_unpair"}
    (the three profiles optional for --noBP) or synth records {"rows", "bp"} of single sequences."""

    def __init__(self, records):
        self.rows = [list(r["rows"]) for r in records]
        prof, pl, pr, pu, off = [], [], [], [], [0]
        for r in records:
            m = hostlib.MData.seq_only(r["rows"])
            p = m.export()["profile"]
            n = p.shape[0]
            prof.append(p)
            if "p_left" in r:
                a, b, c = r["p_left"], r["p_right"], r["p_unpair"]
            elif "bp" in r and len(r["rows"]) == 1:
                a, b, c = pairing_profiles(n, r["bp"][0])
            else:
                a, b, c = np.zeros(n), np.zeros(n), np.ones(n)
            pl.append(np.asarray(a, dtype=np.float32)); pr.append(np.asarray(b, dtype=np.float32))
            pu.append(np.asarray(c, dtype=np.float32))
            off.append(off[-1] + n)
        cat = lambda v, shape: np.ascontiguousarray(np.concatenate(v) if v else np.zeros(shape), dtype=np.float32)
        self.col_off = np.asarray(off, dtype=np.uint32)
        self.profile, self.p_left, self.p_right, self.p_unpair = cat(prof, (0, 5)), cat(pl, 0), cat(pr, 0), cat(pu, 0)

    def __len__(self):
        return len(self.rows)

    def c(self):
        s = BplaSetC(len(self), self.col_off.ctypes.data, self.profile.ctypes.data, self.p_left.ctypes.data,
                     self.p_right.ctypes.data, self.p_unpair.ctypes.data)
        s._owner = self
        return s


def pairs(ctx, params, x, y, xi, yi):
    """k_bpla(x[xi[k]], y[yi[k]]) on the context's device."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    ctx._check(L.lib().stemk_bpla_pairs(ctx.h, C.byref(params), C.byref(cx), C.byref(cy), len(xi), xi.ctypes.data,
                                        yi.ctypes.data, out.ctypes.data))
    return out


def gradients(ctx, params, x, y, xi, yi):
    """BPLAKernel::compute_gradients(x[xi[k]], y[yi[k]], score_table, [alpha, beta, gap, ext], d) on the context's
    device: (values [n], gradients [n, 4] = d/d{alpha, beta, gap, ext}); bpla_kernel.cpp:387-402."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    val, grad = np.zeros(len(xi)), np.zeros((len(xi), 4))
    cx, cy = x.c(), y.c()
    ctx._check(L.lib().stemk_bpla_gradients(ctx.h, C.byref(params), C.byref(cx), C.byref(cy), len(xi), xi.ctypes.data,
                                            yi.ctypes.data, val.ctypes.data, grad.ctypes.data))
    return val, grad


def gradient_matrices(ctx, params, s):
    """The kernel matrix and the four gradient matrices bpla_optimizer builds from compute_gradients over all pairs
    i <= j, mirrored (CalcMatrix, bpla_optimizer.cpp:55-123): (kmat [n, n], gmat [4, n, n])."""
    n = len(s)
    iu = np.triu_indices(n)
    v, g = gradients(ctx, params, s, s, iu[0], iu[1])
    kmat, gmat = np.zeros((n, n)), np.zeros((4, n, n))
    kmat[iu] = v
    kmat[(iu[1], iu[0])] = v
    for l in range(4):
        gmat[l][iu] = g[:, l]
        gmat[l][(iu[1], iu[0])] = g[:, l]
    return kmat, gmat


def gram(ctx, params, s, normalize=False):
    """KernelMatrix::calculate(train, kernel, normalize) for the BPLA kernel (upper triangle evaluated, mirrored)."""
    n = len(s)
    iu = np.triu_indices(n)
    v = pairs(ctx, params, s, s, iu[0], iu[1])
    m = np.zeros((n, n))
    m[iu] = v
    m[(iu[1], iu[0])] = v
    if normalize:
        d = np.diag(m).copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            m = m / np.sqrt(np.outer(d, d))
        np.fill_diagonal(m, 1.0)
    return m
