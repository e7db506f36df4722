"""Naive stem kernel: Python mirror of the stem_kernel/ program's kernel over the C ABI.

    StemKernel<double, BPMat>(use_GU, loop, gap, stack, subst, band=0, ali_bound=0, bp_bound)   stem_kernel/stem_kernel.h:25-52

full_dp only (stem_kernel.cpp:282-351).  Sequences are raw lower-case strings (the program compares characters);
base pairs are either the canonical ones of the text (NormalBasePair / WobbleBasePair) or dense probability tables
(the role of the ViennaRNA-backed BPMatrix class).  Values come from the CUDA library only (stemk_nstem_pairs)."""
import ctypes as C

import numpy as np

from . import _lib as L


class NstemParams(C.Structure):
    """stemk_nstem_params."""
    _fields_ = [("bp_mode", C.c_int32), ("use_gu", C.c_int32), ("loop", C.c_uint32), ("bp_bound", C.c_float),
                ("gap", C.c_double), ("stack", C.c_double), ("subst", C.c_double)]


def make_params(bp_mode=0, use_gu=False, loop=3, gap=0.8, stack=1.0, subst=0.5, bp_bound=0.5):
    """Defaults of stem_kernel/main.cpp:46-64 (parsed as float there); bp_bound: the program's own default (1.0 for
    the canonical classes) makes every kernel value 1, so the mirror defaults to 0.5 = "canonical pairs count"."""
    f = lambda v: float(np.float32(v))
    return NstemParams(int(bp_mode), int(use_gu), int(loop), float(bp_bound), f(gap), f(stack), f(subst))


class NstemSetC(C.Structure):
    """stemk_nstem_set."""
    _fields_ = [("n_seqs", C.c_uint32), ("off", C.c_void_p), ("text", C.c_char_p), ("bp_off", C.c_void_p), ("bp", C.c_void_p)]


def dense_bp(length, bp, th=0.0):
    """Dense L x L float table prob(i, j), 0-based, i < j, from a base-pair list (bi, bj, p) with 1-based positions."""
    t = np.zeros((length, length), dtype=np.float32)
    bi, bj, pp = (np.asarray(a) for a in bp)
    keep = pp >= th
    t[bi[keep] - 1, bj[keep] - 1] = pp[keep]
    return t


class NstemSet:
    """seqs: lower-case strings; tables: optional list of dense L x L float32 probability tables (bp_mode 1)."""

    def __init__(self, seqs, tables=None):
        self.seqs = [s.lower() for s in seqs]
        self.off = np.zeros(len(seqs) + 1, dtype=np.uint32)
        self.off[1:] = np.cumsum([len(s) for s in self.seqs])
        self.text = "".join(self.seqs).encode()
        if tables is not None:
            assert all(t.shape == (len(s), len(s)) for t, s in zip(tables, self.seqs))
            sizes = [len(s) ** 2 for s in self.seqs]
            self.bp_off = np.zeros(len(seqs), dtype=np.uint64)
            self.bp_off[1:] = np.cumsum(sizes[:-1])
            self.bp = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.float32).reshape(-1) for t in tables])
                                           if tables else np.zeros(0), dtype=np.float32)
        else:
            self.bp_off, self.bp = None, None

    def __len__(self):
        return len(self.seqs)

    def c(self):
        s = NstemSetC(len(self), self.off.ctypes.data, self.text, self.bp_off.ctypes.data if self.bp is not None else None,
                      self.bp.ctypes.data if self.bp is not None else None)
        s._owner = self
        return s


def pairs(ctx, params, x, y, xi, yi):
    """k(x[xi[k]], y[yi[k]]) on the context's device."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    ctx._check(L.lib().stemk_nstem_pairs(ctx.h, C.byref(params), C.byref(cx), C.byref(cy), len(xi), xi.ctypes.data,
                                         yi.ctypes.data, out.ctypes.data))
    return out


def pairs_banded(ctx, params, band, x, y, xi, yi):
    """StemKernel(..., band, ali_bound=0)(x[xi[k]], y[yi[k]]) with band > 0: partial_dp (stem_kernel.cpp:113-280)."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    ctx._check(L.lib().stemk_nstem_pairs_banded(ctx.h, C.byref(params), int(band), C.byref(cx), C.byref(cy), len(xi),
                                                xi.ctypes.data, yi.ctypes.data, out.ctypes.data))
    return out


def pack_windows(windows):
    """[(c_low, c_high)] per pair (arrays of lx + 1 entries) -> (win_off, c_low, c_high) as the C ABI takes them."""
    off = np.zeros(len(windows) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(w[0]) for w in windows])
    lo = np.ascontiguousarray(np.concatenate([np.asarray(w[0]) for w in windows]) if windows else np.zeros(0), dtype=np.uint32)
    hi = np.ascontiguousarray(np.concatenate([np.asarray(w[1]) for w in windows]) if windows else np.zeros(0), dtype=np.uint32)
    return off, lo, hi


def pairs_windows(ctx, params, x, y, xi, yi, windows):
    """partial_dp under caller-supplied per-row constraints (the c_low / c_high of StemKernel::alignment_constraints with
    ali_bound > 0, stem_kernel.cpp:14-67): windows[k] = (c_low, c_high) of pair k, lx + 1 entries each."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    off, lo, hi = pack_windows(windows)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    ctx._check(L.lib().stemk_nstem_pairs_windows(ctx.h, C.byref(params), C.byref(cx), C.byref(cy), len(xi), xi.ctypes.data,
                                                 yi.ctypes.data, off.ctypes.data, lo.ctypes.data, hi.ctypes.data, out.ctypes.data))
    return out
