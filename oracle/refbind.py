"""ctypes binding of oracle/_ref/libstemk_ref*.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The shared objects are the UNMODIFIED reference sources (see oracle/Makefile and
oracle/ref_harness*.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# STEMK_REF_VARIANT=o3 selects the courtesy build of the same sources (-O3 -march=x86-64-v3)
REF_DIR = os.path.join(_HERE, "_ref", os.environ.get("STEMK_REF_VARIANT", ""))

# kernel kinds of ref_harness.cpp (RefKind)
SI_STEM, SU_STEM, SI_STEM_STR, SU_STEM_STR, LSU_STEM, LSU_STR, LSU_STEM_STR, STR_SUBST, STR_SIMPLE = range(9)


class RefParams(C.Structure):
    _fields_ = [("kind", C.c_int), ("loop_gap", C.c_double), ("beta", C.c_double), ("stack", C.c_double),
                ("covar", C.c_double), ("gap", C.c_double), ("alpha", C.c_double), ("match", C.c_double),
                ("mismatch", C.c_double), ("len_band", C.c_uint)]


def available():
    return os.path.exists(os.path.join(REF_DIR, "libstemk_ref.so"))


_libs = {}


def _lib(name):
    if name not in _libs:
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        _libs[name] = C.CDLL(path)
    return _libs[name]


# which build of the reference harness lib() hands out: "libstemk_ref.so" (the checker proper) or
# "libstemk_ref_binding.so" (the same translation units + the KernelMatrix binding of INTEGRATION.md linked against the
# product library; tests/test_ref_binding.py switches to it with use_library()).  MData / kernel handles belong to the
# library that made them.
_DEFAULT = "libstemk_ref.so"


def use_library(name):
    global _DEFAULT
    _DEFAULT = name


def lib():
    L = _lib(_DEFAULT)
    if getattr(L, "_typed", False):
        return L
    vp, ci, cu, cd, cf = C.c_void_p, C.c_int, C.c_uint, C.c_double, C.c_float
    P = C.POINTER
    L.ref_mdata_new.restype = vp
    L.ref_mdata_new.argtypes = [ci, P(C.c_char_p), P(P(cd)), cf]
    L.ref_mdata_new_seqonly.restype = vp
    L.ref_mdata_new_seqonly.argtypes = [ci, P(C.c_char_p)]
    L.ref_mdata_from_flat.restype = vp
    L.ref_mdata_from_flat.argtypes = [ci] + [vp] * 12 + [ci, vp, ci, P(C.c_char_p)]
    L.ref_mdata_free.argtypes = [vp]
    L.ref_mdata_counts.argtypes = [vp] + [P(ci)] * 6
    L.ref_mdata_dump.argtypes = [vp] * 19
    L.ref_kernel_new.restype = vp
    L.ref_kernel_new.argtypes = [P(RefParams)]
    L.ref_kernel_free.argtypes = [vp]
    L.ref_kernel_pair.restype = cd
    L.ref_kernel_pair.argtypes = [vp, vp, vp]
    L.ref_gram.restype = cd
    L.ref_gram.argtypes = [vp, ci, P(vp), vp, ci, cu, vp, vp, C.c_long, P(C.c_long)]
    L.ref_cross.restype = cd
    L.ref_cross.argtypes = [vp, ci, P(vp), ci, P(vp), ci, ci, cu, vp, vp]
    L.ref_row.restype = cd
    L.ref_row.argtypes = [vp, vp, ci, P(vp), vp, ci, cu, vp, vp]
    L.ref_diag.restype = cd
    L.ref_diag.argtypes = [vp, ci, P(vp), vp, ci, cu, vp]
    L.ref_pairs_timed.restype = cd
    L.ref_pairs_timed.argtypes = [vp, ci, P(vp), ci, vp, vp, cu, vp]
    if hasattr(L, "refbind_gram"):   # the compiled INTEGRATION.md binding (oracle/ref_binding_test.cpp)
        L.refbind_gram.restype = ci
        L.refbind_gram.argtypes = [vp, ci, P(vp), vp, ci, ci, vp, vp, C.c_long, P(C.c_long), vp, C.c_long]
        L.refbind_cross.restype = ci
        L.refbind_cross.argtypes = [vp, ci, P(vp), ci, P(vp), ci, ci, ci, vp, vp, vp, C.c_long]
        L.refbind_diag.restype = ci
        L.refbind_diag.argtypes = [vp, ci, P(vp), vp, ci, ci, vp, vp, C.c_long]
    L._typed = True
    return L


def _rows(rows):
    arr = (C.c_char_p * len(rows))(*[r.encode() if isinstance(r, str) else r for r in rows])
    return arr


def dense_bp(L, bi, bj, bp):
    """(L+1)x(L+1) row-major 1-based matrix from a sparse pair list (1-based i<j)."""
    m = np.zeros((L + 1, L + 1), dtype=np.float64)
    if len(bi):
        m[np.asarray(bi), np.asarray(bj)] = np.asarray(bp)
    return m


class RefMData:
    """One reference MData (stem_kernel_lite/data.h:26-53) built by the reference's constructor."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference MData construction failed")
        self.h = handle

    @classmethod
    def build(cls, rows, bp_sparse_rows, th):
        """rows: aligned strings; bp_sparse_rows: per row (bi, bj, bp) over the UNGAPPED row, 1-based."""
        mats = []
        for r, (bi, bj, bp) in zip(rows, bp_sparse_rows):
            Lr = sum(1 for c in r if c != "-")
            mats.append(np.ascontiguousarray(dense_bp(Lr, bi, bj, bp)))
        ptrs = (C.POINTER(C.c_double) * len(rows))(*[m.ctypes.data_as(C.POINTER(C.c_double)) for m in mats])
        return cls(lib().ref_mdata_new(len(rows), _rows(rows), ptrs, C.c_float(th)))

    @classmethod
    def seq_only(cls, rows):
        return cls(lib().ref_mdata_new_seqonly(len(rows), _rows(rows)))

    @classmethod
    def from_flat(cls, f, rows):
        """f: dict with the arrays of dump() (edge_ppos/edge_cpos included)."""
        a = lambda k, dt: np.ascontiguousarray(f[k], dtype=dt)
        keep = [a("first", np.uint32), a("last", np.uint32), a("weight", np.float32), a("edge_off", np.uint32),
                a("edge_to", np.uint32), a("edge_ppos", np.uint32), a("edge_cpos", np.uint32),
                a("edge_w", np.float32), a("bpf_off", np.uint32), a("bpf_a", np.uint8), a("bpf_b", np.uint8),
                a("bpf_f", np.float32)]
        sw = a("seq_weight", np.float32)
        L = len(rows[0])
        h = lib().ref_mdata_from_flat(len(keep[0]), *[k.ctypes.data for k in keep], L,
                                      sw.ctypes.data if len(sw) else None, len(rows), _rows(rows))
        return cls(h)

    def __del__(self):
        try:
            if self.h:
                lib().ref_mdata_free(self.h)
                self.h = None
        except Exception:
            pass

    def dump(self):
        n = [C.c_int() for _ in range(6)]
        lib().ref_mdata_counts(self.h, *[C.byref(x) for x in n])
        nn, ne, nb, nr, L, nw = [x.value for x in n]
        u32, f32, u8 = np.uint32, np.float32, np.uint8
        d = dict(first=np.zeros(nn, u32), last=np.zeros(nn, u32), weight=np.zeros(nn, f32),
                 edge_off=np.zeros(nn + 1, u32), edge_to=np.zeros(ne, u32), edge_gaps=np.zeros(ne, u32),
                 edge_w=np.zeros(ne, f32), edge_ppos=np.zeros(2 * ne, u32), edge_cpos=np.zeros(2 * ne, u32),
                 bpf_off=np.zeros(nn + 1, u32), bpf_a=np.zeros(nb, u8), bpf_b=np.zeros(nb, u8),
                 bpf_f=np.zeros(nb, f32), root=np.zeros(nr, u32), max_pa=np.zeros(nn, u32),
                 profile=np.zeros((L, 5), f32), n_seqs=np.zeros(1, f32), seq_weight=np.zeros(nw, f32))
        order = ["first", "last", "weight", "edge_off", "edge_to", "edge_gaps", "edge_w", "edge_ppos", "edge_cpos",
                 "bpf_off", "bpf_a", "bpf_b", "bpf_f", "root", "max_pa", "profile", "n_seqs", "seq_weight"]
        lib().ref_mdata_dump(self.h, *[d[k].ctypes.data for k in order])
        d["n_seqs"] = float(d["n_seqs"][0])
        return d


class RefKernel:
    def __init__(self, kind, loop_gap=0.2, beta=0.3, stack=1.3, covar=0.8, gap=0.8, alpha=0.2, match=1.0,
                 mismatch=0.8, len_band=10):
        self.p = RefParams(kind, loop_gap, beta, stack, covar, gap, alpha, match, mismatch, len_band)
        self.h = lib().ref_kernel_new(C.byref(self.p))

    def __del__(self):
        try:
            lib().ref_kernel_free(self.h)
        except Exception:
            pass

    @staticmethod
    def _hs(ds):
        return (C.c_void_p * len(ds))(*[d.h for d in ds])

    def pair(self, x, y):
        return lib().ref_kernel_pair(self.h, x.h, y.h)

    def gram(self, ds, normalize=False, n_th=1, labels=None, want_text=False):
        n = len(ds)
        out = np.zeros((n, n))
        lab = np.ascontiguousarray(labels, dtype=np.int32) if labels is not None else None
        text = C.create_string_buffer(64 + n * (n + 2) * 24) if want_text else None
        tl = C.c_long(0)
        secs = lib().ref_gram(self.h, n, self._hs(ds), lab.ctypes.data if lab is not None else None, int(normalize),
                              n_th, out.ctypes.data, text, len(text) if text else 0, C.byref(tl))
        if want_text:
            return out, secs, text.raw[:tl.value].decode()
        return out, secs

    def cross(self, test, train, norm_test=False, normalize=False, n_th=1):
        out = np.zeros((len(test), len(train)))
        selfv = np.zeros(len(test))
        secs = lib().ref_cross(self.h, len(test), self._hs(test), len(train), self._hs(train), int(norm_test),
                               int(normalize), n_th, out.ctypes.data, selfv.ctypes.data)
        return out, selfv, secs

    def row(self, test, train, sv_index=(), n_th=1, want_self=True, init=0.0):
        out = np.full(len(train), init, dtype=np.float64)
        sv = np.ascontiguousarray(sv_index, dtype=np.uint32)
        selfv = C.c_double(0)
        secs = lib().ref_row(self.h, test.h, len(train), self._hs(train), sv.ctypes.data, len(sv), n_th,
                             out.ctypes.data, C.byref(selfv) if want_self else None)
        return out, selfv.value, secs

    def diag(self, train, sv_index=(), n_th=1, init=0.0):
        out = np.full(len(train), init, dtype=np.float64)
        sv = np.ascontiguousarray(sv_index, dtype=np.uint32)
        secs = lib().ref_diag(self.h, len(train), self._hs(train), sv.ctypes.data, len(sv), n_th, out.ctypes.data)
        return out, secs

    # ---- the same KernelMatrix members through the compiled binding (GpuBound<K>, oracle/ref_binding_test.cpp)
    @staticmethod
    def _bound(rc, err):
        if rc != 0:
            raise RuntimeError("reference-side binding: " + err.value.decode(errors="replace"))

    def bound_gram(self, ds, normalize=False, labels=None, device=0):
        n = len(ds)
        out = np.zeros((n, n))
        lab = np.ascontiguousarray(labels, dtype=np.int32) if labels is not None else None
        text = C.create_string_buffer(64 + n * (n + 2) * 24)
        tl, err = C.c_long(0), C.create_string_buffer(512)
        self._bound(lib().refbind_gram(self.h, n, self._hs(ds), lab.ctypes.data if lab is not None else None, int(normalize),
                                       device, out.ctypes.data, text, len(text), C.byref(tl), err, len(err)), err)
        return out, text.raw[:tl.value].decode()

    def bound_cross(self, test, train, norm_test=False, normalize=False, device=0):
        out = np.zeros((len(test), len(train)))
        selfv = np.zeros(len(test))
        err = C.create_string_buffer(512)
        self._bound(lib().refbind_cross(self.h, len(test), self._hs(test), len(train), self._hs(train), int(norm_test),
                                        int(normalize), device, out.ctypes.data, selfv.ctypes.data, err, len(err)), err)
        return out, selfv

    def bound_diag(self, train, sv_index=(), device=0, init=0.0):
        out = np.full(len(train), init, dtype=np.float64)
        sv = np.ascontiguousarray(sv_index, dtype=np.uint32)
        err = C.create_string_buffer(512)
        self._bound(lib().refbind_diag(self.h, len(train), self._hs(train), sv.ctypes.data, len(sv), device, out.ctypes.data,
                                       err, len(err)), err)
        return out

    def pairs_timed(self, ds, pi, pj, n_th=1):
        pi = np.ascontiguousarray(pi, dtype=np.int32)
        pj = np.ascontiguousarray(pj, dtype=np.int32)
        out = np.zeros(len(pi))
        secs = lib().ref_pairs_timed(self.h, len(ds), self._hs(ds), len(pi), pi.ctypes.data, pj.ctypes.data, n_th,
                                     out.ctypes.data)
        return out, secs


# ---- naive string kernel (string_kernel/string_kernel.cpp) ----
def naive_lib():
    L = _lib("libstemk_ref_naive.so")
    if not getattr(L, "_typed", False):
        L.refn_pair.restype = C.c_double
        L.refn_pair.argtypes = [C.c_float, C.c_char_p, C.c_char_p]
        L.refn_gram.restype = C.c_double
        L.refn_gram.argtypes = [C.c_float, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_uint, C.c_void_p]
        L.refn_cross.restype = C.c_double
        L.refn_cross.argtypes = [C.c_float, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p),
                                 C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p]
        L._typed = True
    return L


def naive_pair(gap, x, y):
    return naive_lib().refn_pair(gap, x.encode(), y.encode())


def naive_gram(gap, seqs, normalize=False, n_th=1):
    n = len(seqs)
    out = np.zeros((n, n))
    secs = naive_lib().refn_gram(gap, n, _rows(seqs), int(normalize), n_th, out.ctypes.data)
    return out, secs


# ---- vendored LIBSVM ----
def svm_lib():
    L = _lib("libstemk_ref_svm.so")
    if not getattr(L, "_typed", False):
        L.refsvm_cv.restype = C.c_int
        L.refsvm_cv.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_uint, C.c_void_p]
        L.refsvm_train_predict.restype = C.c_int
        L.refsvm_train_predict.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p,
                                           C.c_void_p]
        L._typed = True
    return L


def svm_cv(K, y, C_=1.0, nr_fold=5, seed=1):
    K = np.ascontiguousarray(K, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    target = np.zeros(len(y))
    rc = svm_lib().refsvm_cv(len(y), K.ctypes.data, y.ctypes.data, C_, nr_fold, seed, target.ctypes.data)
    assert rc == 0
    return target


def svm_train_predict(Ktrain, y, Ktest, C_=1.0):
    Ktrain = np.ascontiguousarray(Ktrain, dtype=np.float64)
    Ktest = np.ascontiguousarray(Ktest, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    pred = np.zeros(Ktest.shape[0])
    rc = svm_lib().refsvm_train_predict(len(y), Ktrain.ctypes.data, y.ctypes.data, C_, Ktest.shape[0],
                                        Ktest.ctypes.data, pred.ctypes.data)
    assert rc == 0
    return pred


# ---- the reference BPLA kernel (bpla_kernel/bpla_kernel.cpp behind oracle/ref_harness_bpla.cpp) ----
def bpla_pairs(params, x, y, xi, yi):
    """params / x / y: stem_kernel_b200.bpla.BplaParams / BplaSet; the reference's own Data objects are rebuilt from
    the rows and the three base-pairing profiles."""
    Lb = _lib("libstemk_ref_bpla.so")
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))

    def rows_of(s):
        off = np.zeros(len(s) + 1, dtype=np.uint32)
        flat = []
        for k, r in enumerate(s.rows):
            flat.extend(r)
            off[k + 1] = len(flat)
        return off, (C.c_char_p * max(1, len(flat)))(*[t.encode() for t in flat])

    ox, rx = rows_of(x)
    oy, ry = rows_of(y)
    table = np.array(list(params.score), dtype=np.float64)
    vp = C.c_void_p
    Lb.refbpla_pairs.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, vp,
                                 C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_size_t, vp, vp, vp]
    rc = Lb.refbpla_pairs(params.no_bp, params.sw, params.gap, params.ext, params.alpha, params.beta, table.ctypes.data,
                          len(x), ox.ctypes.data, rx, x.col_off.ctypes.data, x.p_left.ctypes.data, x.p_right.ctypes.data,
                          x.p_unpair.ctypes.data, len(y), oy.ctypes.data, ry, y.col_off.ctypes.data, y.p_left.ctypes.data,
                          y.p_right.ctypes.data, y.p_unpair.ctypes.data, len(xi), xi.ctypes.data, yi.ctypes.data,
                          out.ctypes.data)
    assert rc == 0
    return out


def bpla_gradients(params, x, y, xi, yi):
    """BPLAKernel<double,MData>::compute_gradients of the compiled reference: (values [n], gradients [n, 4])."""
    Lb = _lib("libstemk_ref_bpla.so")
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    val, grad = np.zeros(len(xi)), np.zeros((len(xi), 4))

    def rows_of(s):
        off = np.zeros(len(s) + 1, dtype=np.uint32)
        flat = []
        for k, r in enumerate(s.rows):
            flat.extend(r)
            off[k + 1] = len(flat)
        return off, (C.c_char_p * max(1, len(flat)))(*[t.encode() for t in flat])

    ox, rx = rows_of(x)
    oy, ry = rows_of(y)
    table = np.array(list(params.score), dtype=np.float64)
    vp = C.c_void_p
    Lb.refbpla_gradients.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, vp,
                                     C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_size_t, vp, vp, vp, vp]
    rc = Lb.refbpla_gradients(params.gap, params.ext, params.alpha, params.beta, table.ctypes.data,
                              len(x), ox.ctypes.data, rx, x.col_off.ctypes.data, x.p_left.ctypes.data, x.p_right.ctypes.data,
                              x.p_unpair.ctypes.data, len(y), oy.ctypes.data, ry, y.col_off.ctypes.data, y.p_left.ctypes.data,
                              y.p_right.ctypes.data, y.p_unpair.ctypes.data, len(xi), xi.ctypes.data, yi.ctypes.data,
                              val.ctypes.data, grad.ctypes.data)
    assert rc == 0
    return val, grad


# ---- the reference naive stem kernel (stem_kernel/stem_kernel.cpp behind oracle/ref_harness_nstem.cpp) ----
def nstem_pairs(params, x, y, xi, yi, band=0, ali_bound=0.0):
    """params / x / y: stem_kernel_b200.nstem.NstemParams / NstemSet."""
    Ln = _lib("libstemk_ref_nstem.so")
    xi = np.ascontiguousarray(xi, dtype=np.uint32)
    yi = np.ascontiguousarray(yi, dtype=np.uint32)
    out = np.zeros(len(xi))
    vp = C.c_void_p
    Ln.refnstem_pairs.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_uint, C.c_float, C.c_float,
                                  C.c_int, vp, C.c_char_p, vp, vp, C.c_int, vp, C.c_char_p, vp, vp, C.c_size_t, vp, vp, vp]
    d = lambda a: a.ctypes.data if a is not None else None
    rc = Ln.refnstem_pairs(params.bp_mode, params.use_gu, params.loop, params.gap, params.stack, params.subst, band, ali_bound,
                           params.bp_bound, len(x), d(x.off), x.text, d(x.bp_off), d(x.bp), len(y), d(y.off), y.text,
                           d(y.bp_off), d(y.bp), len(xi), xi.ctypes.data, yi.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def nstem_windows(x_seq, y_seq, band=0, ali_bound=0.0):
    """(c_low, c_high), lx + 1 entries each: StemKernel::alignment_constraints of the reference (stem_kernel.cpp:14-83) --
    the pair-HMM constraints when ali_bound > 0 (narrowed by the band), the band alone otherwise."""
    Ln = _lib("libstemk_ref_nstem.so")
    Ln.refnstem_windows.argtypes = [C.c_uint, C.c_float, C.c_char_p, C.c_uint, C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p]
    xs, ys = x_seq.lower().encode(), y_seq.lower().encode()
    lo, hi = np.zeros(len(xs) + 1, dtype=np.uint32), np.zeros(len(xs) + 1, dtype=np.uint32)
    rc = Ln.refnstem_windows(int(band), float(ali_bound), xs, len(xs), ys, len(ys), lo.ctypes.data, hi.ctypes.data)
    assert rc == 0
    return lo, hi
