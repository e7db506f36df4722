"""ctypes binding of oracle/libstemk_oracle.so (plain-C restatement) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "libstemk_oracle.so")


class Params(C.Structure):
    """stemk_params of include/stemk.h."""
    _fields_ = [("kind", C.c_int32), ("len_band", C.c_uint32), ("loop_gap", C.c_double), ("beta", C.c_double),
                ("stack", C.c_double), ("covar", C.c_double), ("gap", C.c_double), ("alpha", C.c_double),
                ("match", C.c_double), ("mismatch", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError(f"{SO} is not built: run `make -C oracle oracle`")
        L = C.CDLL(SO)
        vp = C.c_void_p
        L.oracle_pair.restype = C.c_double
        L.oracle_pair.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32]
        L.oracle_pairs.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp]
        L.oracle_gram.argtypes = [vp, vp, C.c_int, vp]
        L.oracle_diag.argtypes = [vp, vp, vp, C.c_uint32, vp]
        L.oracle_cross.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_int, vp, vp]
        L.oracle_print.restype = C.c_long
        L.oracle_print.argtypes = [vp, C.c_size_t, C.c_size_t, vp, vp, C.c_long]
        L.oracle_pair_cost.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp]
        L.oracle_bpla_pairs.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp]
        L.oracle_bpla_gradients.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp, vp]
        L.oracle_nstem_pairs.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp]
        L.oracle_nstem_pairs_banded.argtypes = [vp, C.c_uint, vp, vp, C.c_size_t, vp, vp, vp]
        L.oracle_nstem_pairs_windows.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]
        L.oracle_fold_bpp.argtypes = [vp, C.c_char_p, C.c_uint32, vp, vp, vp]
        L.oracle_fold_model_default.argtypes = [vp]
        L.oracle_fold_default_scale.restype = C.c_double
        L.oracle_fold_default_scale.argtypes = [C.c_double]
        _lib = L
    return _lib


def _p(x):
    return C.byref(x)


def pair(params, dx, xi, dy, yi):
    return lib().oracle_pair(_p(params), _p(dx), xi, _p(dy), yi)


def pairs(params, dx, dy, xi, yi):
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    lib().oracle_pairs(_p(params), _p(dx), _p(dy), len(xi), xi.ctypes.data, yi.ctypes.data, out.ctypes.data)
    return out


def gram(params, d, normalize=False):
    out = np.zeros((d.n_seqs, d.n_seqs))
    lib().oracle_gram(_p(params), _p(d), int(normalize), out.ctypes.data)
    return out


def diag(params, d, sv_index=(), init=0.0):
    out = np.full(d.n_seqs, init, dtype=np.float64)
    sv = np.ascontiguousarray(sv_index, dtype=np.uint32)
    lib().oracle_diag(_p(params), _p(d), sv.ctypes.data, len(sv), out.ctypes.data)
    return out


def cross(params, dt, ds, sv_index=(), normalize=False, want_self=True, init=0.0):
    out = np.full((dt.n_seqs, ds.n_seqs), init, dtype=np.float64)
    selfv = np.zeros(dt.n_seqs)
    sv = np.ascontiguousarray(sv_index, dtype=np.uint32)
    lib().oracle_cross(_p(params), _p(dt), _p(ds), sv.ctypes.data, len(sv), int(normalize), out.ctypes.data,
                       selfv.ctypes.data if want_self else None)
    return out, selfv


def print_matrix(m, labels=None):
    m = np.ascontiguousarray(m, dtype=np.float64)
    lab = np.ascontiguousarray(labels, dtype=np.int32) if labels is not None else None
    cap = 64 + m.shape[0] * (m.shape[1] + 2) * 28
    buf = C.create_string_buffer(cap)
    n = lib().oracle_print(m.ctypes.data, m.shape[0], m.shape[1], lab.ctypes.data if lab is not None else None, buf,
                           cap)
    return buf.raw[:n].decode()


def pair_cost(params, dx, xi, dy, yi):
    c, f = C.c_double(), C.c_double()
    lib().oracle_pair_cost(_p(params), _p(dx), xi, _p(dy), yi, C.byref(c), C.byref(f))
    return c.value, f.value


def bpla_pairs(params, x, y, xi, yi):
    """Restated BPLA kernel; params / x / y: stem_kernel_b200.bpla.BplaParams / BplaSet (same C structs)."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    lib().oracle_bpla_pairs(_p(params), _p(cx), _p(cy), len(xi), xi.ctypes.data, yi.ctypes.data, out.ctypes.data)
    return out


def bpla_gradients(params, x, y, xi, yi):
    """Restated BPLAKernel::compute_gradients: (values [n], gradients [n, 4] = d/d{alpha, beta, gap, ext})."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    val, grad = np.zeros(len(xi)), np.zeros((len(xi), 4))
    cx, cy = x.c(), y.c()
    lib().oracle_bpla_gradients(_p(params), _p(cx), _p(cy), len(xi), xi.ctypes.data, yi.ctypes.data, val.ctypes.data,
                                grad.ctypes.data)
    return val, grad


def nstem_pairs(params, x, y, xi, yi):
    """Restated naive stem kernel; params / x / y: stem_kernel_b200.nstem.NstemParams / NstemSet."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    lib().oracle_nstem_pairs(_p(params), _p(cx), _p(cy), len(xi), xi.ctypes.data, yi.ctypes.data, out.ctypes.data)
    return out


def nstem_pairs_banded(params, band, x, y, xi, yi):
    """Restated StemKernel::partial_dp with the band-only constraints (band > 0; ali_bound = 0)."""
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    lib().oracle_nstem_pairs_banded(_p(params), int(band), _p(cx), _p(cy), len(xi), xi.ctypes.data, yi.ctypes.data, out.ctypes.data)
    return out


def nstem_pairs_windows(params, x, y, xi, yi, windows):
    """Restated partial_dp under caller-supplied per-row windows; windows[k] = (c_low, c_high), lx + 1 entries each."""
    from stem_kernel_b200.nstem import pack_windows
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    off, lo, hi = pack_windows(windows)
    out = np.zeros(len(xi))
    cx, cy = x.c(), y.c()
    lib().oracle_nstem_pairs_windows(_p(params), _p(cx), _p(cy), len(xi), xi.ctypes.data, yi.ctypes.data, off.ctypes.data,
                                     lo.ctypes.data, hi.ctypes.data, out.ctypes.data)
    return out


class FoldModel(C.Structure):
    """stemk_fold_model of include/stemk.h."""
    _fields_ = [("temperature", C.c_double), ("pf_scale", C.c_double), ("no_gu", C.c_int32), ("pad_", C.c_int32),
                ("stack", C.c_double * 8 * 8), ("hairpin", C.c_double * 31), ("bulge", C.c_double * 31),
                ("interior", C.c_double * 31), ("lxc", C.c_double), ("mismatch_h", C.c_double * 5 * 5 * 8),
                ("mismatch_i", C.c_double * 5 * 5 * 8), ("dangle5", C.c_double * 5 * 8), ("dangle3", C.c_double * 5 * 8),
                ("ninio", C.c_double), ("max_ninio", C.c_double), ("terminal_au", C.c_double), ("ml_closing", C.c_double),
                ("ml_intern", C.c_double * 8), ("ml_base", C.c_double)]


def fold_model_default():
    m = FoldModel()
    lib().oracle_fold_model_default(_p(m))
    return m


def fold_bpp(model, seq):
    """(dense (L+1)x(L+1) table with P(i,j) at [i,j], 1-based i<j; ensemble free energy; unpaired[L]) of one sequence
    (oracle/stemk_fold_oracle.c: McCaskill restated; parity with ViennaRNA unpinned)."""
    n = len(seq)
    dense = np.zeros((n + 1, n + 1))
    unp = np.zeros(n)
    ens = C.c_double(0.0)
    rc = lib().oracle_fold_bpp(_p(model), seq.encode(), n, dense.ctypes.data, _p(ens), unp.ctypes.data)
    if rc != 0:
        raise FloatingPointError("partition function out of range under this pf_scale")
    return dense, ens.value, unp
