"""Python mirror of the reference's kernel classes and KernelMatrix, over the C ABI (include/stemk.h).

    SuStemKernel / SiStemKernel / SuStemStrKernel / ...   stem_kernel_lite/def_kernel.h:12-192
    StringKernel                                           stem_kernel_lite/string_kernel.h:8-30
    NaiveStringKernel                                      string_kernel/string_kernel.h:8-22
    KernelMatrix.calculate / calculate_test / diagonal / print
                                                           common/kernel_matrix.h:67-107, .cpp:485-770

A kernel object is immutable (like the reference's); KernelMatrix owns an n x n (or n_test x n_train)
numpy matrix and the labels, and prints LIBSVM "precomputed kernel" text.  All arithmetic happens in the
CUDA library; this module only marshals arrays."""
import ctypes as C

import numpy as np

from . import _lib as L
from .hostlib import MData, SeqSet


class StemkError(RuntimeError):
    pass


class Context:
    """stemk_ctx: one device + one kernel object."""

    def __init__(self, params, device=0):
        self.params = params
        self.h = C.c_void_p()
        rc = L.lib().stemk_create(C.byref(self.h), C.byref(params), device)
        if rc != L.OK:
            raise StemkError(f"stemk_create failed ({rc}): {L.lib().stemk_last_error(None).decode()}")
        self.device = device

    def _check(self, rc):
        if rc != L.OK:
            raise StemkError(f"stemk error {rc}: {L.lib().stemk_last_error(self.h).decode()}")

    def set_option(self, option, value=1):
        """stemk_set_option: L.OPT_FORCE_GENERAL (general stem kernel for every pair), L.OPT_FORCE_UNSTAGED (its any-size variant), L.OPT_TIMING."""
        self._check(L.lib().stemk_set_option(self.h, int(option), int(value)))
        return self

    def close(self):
        if self.h:
            L.lib().stemk_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, mdatas):
        return DeviceSet(self, mdatas)

    # ---- KernelMatrix entry points
    def gram(self, dset, normalize=False):
        n = len(dset)
        out = np.zeros((n, n))
        self._check(L.lib().stemk_gram(self.h, dset.h, int(normalize), out.ctypes.data))
        return out

    def cross(self, test, train, sv_index=None, normalize=False, want_self=True, init=0.0):
        out = np.full((len(test), len(train)), init, dtype=np.float64)
        selfv = np.zeros(len(test))
        sv = np.ascontiguousarray(sv_index if sv_index is not None else [], dtype=np.uint32)
        self._check(L.lib().stemk_cross(self.h, test.h, train.h, sv.ctypes.data if len(sv) else None, len(sv),
                                        int(normalize), out.ctypes.data, selfv.ctypes.data if want_self else None))
        return out, selfv

    def diag(self, train, sv_index=None, init=0.0):
        out = np.full(len(train), init, dtype=np.float64)
        sv = np.ascontiguousarray(sv_index if sv_index is not None else [], dtype=np.uint32)
        self._check(L.lib().stemk_diag(self.h, train.h, sv.ctypes.data if len(sv) else None, len(sv), out.ctypes.data))
        return out

    def pairs(self, x, y, xi, yi):
        xi = np.ascontiguousarray(xi, dtype=np.uint32)
        yi = np.ascontiguousarray(yi, dtype=np.uint32)
        out = np.zeros(len(xi))
        self._check(L.lib().stemk_pairs(self.h, x.h, y.h, len(xi), xi.ctypes.data, yi.ctypes.data, out.ctypes.data))
        return out

    def pairs_device(self, x, y, n_pairs, d_xi, d_yi, d_out, stream=None):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on `stream` (cudaStream_t as int)."""
        self._check(L.lib().stemk_pairs_device(self.h, x.h, y.h, n_pairs, d_xi, d_yi, d_out, stream))

    def assemble_device(self, n_pairs, d_xi, d_yi, d_vals, n, normalize, d_matrix, stream=None):
        """Scatter gathered pair values into the n x n device matrix (+ mirror, + normalisation)."""
        self._check(L.lib().stemk_assemble_device(self.h, n_pairs, d_xi, d_yi, d_vals, n, int(normalize), d_matrix,
                                                  stream))

    def pair_cost(self, x, y, xi, yi):
        xi = np.ascontiguousarray(xi, dtype=np.uint32)
        yi = np.ascontiguousarray(yi, dtype=np.uint32)
        cells, flops = np.zeros(len(xi)), np.zeros(len(xi))
        self._check(L.lib().stemk_pair_cost(self.h, x.h, y.h, len(xi), xi.ctypes.data, yi.ctypes.data,
                                            cells.ctypes.data, flops.ctypes.data))
        return cells, flops

    def stats_reset(self):
        L.lib().stemk_stats_reset(self.h)

    def stats(self):
        n, a, b = C.c_uint64(), C.c_double(), C.c_double()
        L.lib().stemk_stats_get(self.h, C.byref(n), C.byref(a), C.byref(b))
        return dict(launches=n.value, stem_ms=a.value, string_ms=b.value)

    def fp64_peak(self, seconds=0.5):
        t = C.c_double()
        self._check(L.lib().stemk_fp64_peak(self.h, seconds, C.byref(t)))
        return t.value


class DeviceSet:
    """stemk_set: a flattened, compiled and uploaded list of MData records."""

    def __init__(self, ctx, mdatas):
        self.ctx = ctx
        self.h = C.c_void_p()
        if mdatas is None:      # adopted handle (clone / import)
            self.n = 0
            return
        if isinstance(mdatas, SeqSet):
            flat = mdatas
        else:
            flat = SeqSet(mdatas)
        desc = flat.desc()
        ctx._check(L.lib().stemk_upload(ctx.h, C.byref(desc), C.byref(self.h)))
        self.n = len(flat)

    @classmethod
    def _adopt(cls, ctx, handle):
        s = cls(ctx, None)
        s.h = handle
        s.n = int(L.lib().stemk_set_size(handle))
        return s

    def clone_to(self, ctx):
        """stemk_set_clone: the same set on the device of another context (device-to-device copy, no recompilation)."""
        h = C.c_void_p()
        ctx._check(L.lib().stemk_set_clone(ctx.h, self.h, C.byref(h)))
        return DeviceSet._adopt(ctx, h)

    def export_bytes(self):
        return int(L.lib().stemk_set_export_bytes(self.h))

    def export_to(self, d_ptr, stream=None):
        """stemk_set_export into a DEVICE buffer of export_bytes() bytes (e.g. a torch uint8 tensor's data_ptr())."""
        self.ctx._check(L.lib().stemk_set_export(self.ctx.h, self.h, d_ptr, stream))

    @classmethod
    def import_from(cls, ctx, d_ptr, nbytes):
        """stemk_set_import: a set exported on another rank, after the caller's broadcast brought its bytes here."""
        h = C.c_void_p()
        ctx._check(L.lib().stemk_set_import(ctx.h, d_ptr, nbytes, C.byref(h)))
        return cls._adopt(ctx, h)

    def __len__(self):
        return self.n

    def stats(self):
        """Per-record (#V, #E, L) as the reference counts them (leaves and leaf edges included)."""
        v, e, l = (np.zeros(self.n, dtype=np.uint32) for _ in range(3))
        L.lib().stemk_set_stats(self.h, v.ctypes.data, e.ctypes.data, l.ctypes.data)
        return v, e, l

    def device_bytes(self):
        return int(L.lib().stemk_set_device_bytes(self.h))

    def free(self):
        if self.h and self.ctx.h:
            L.lib().stemk_set_free(self.ctx.h, self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def upload_multi(ctxs, mdatas):
    """stemk_upload_multi: compile once, one set per context/device (sets[d] on ctxs[d])."""
    flat = mdatas if isinstance(mdatas, SeqSet) else SeqSet(mdatas)
    desc = flat.desc()
    hc = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    hs = (C.c_void_p * len(ctxs))()
    ctxs[0]._check(L.lib().stemk_upload_multi(hc, len(ctxs), C.byref(desc), hs))
    return [DeviceSet._adopt(c, C.c_void_p(h)) for c, h in zip(ctxs, hs)]


def gram_multi(ctxs, sets, normalize=False):
    """stemk_gram_multi: one Gram matrix over the devices of `ctxs` (one host thread, one context per device)."""
    n = len(sets[0])
    out = np.zeros((n, n))
    hc = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    hs = (C.c_void_p * len(ctxs))(*[s.h for s in sets])
    ctxs[0]._check(L.lib().stemk_gram_multi(hc, hs, len(ctxs), int(normalize), out.ctypes.data))
    return out


# ---------------------------------------------------------------- kernel classes (def_kernel.h)
class _Kernel:
    kind = None

    def __init__(self, device=0, **kw):
        self.params = L.make_params(self.kind, **kw)
        self.ctx = Context(self.params, device)

    def __call__(self, x, y):
        """value_type operator()(const Data&, const Data&) const -- one pair (x, y are MData)."""
        sx, sy = self.ctx.upload([x]), self.ctx.upload([y])
        return float(self.ctx.pairs(sx, sy, [0], [0])[0])


class SiStemKernel(_Kernel):
    kind = L.SI_STEM

    def __init__(self, loop_gap=0.2, stack=1.3, covar=0.8, len_band=10, device=0):
        super().__init__(device, loop_gap=loop_gap, stack=stack, covar=covar, len_band=len_band)


class SuStemKernel(_Kernel):
    kind = L.SU_STEM

    def __init__(self, loop_gap=0.2, beta=0.3, len_band=10, device=0):
        super().__init__(device, loop_gap=loop_gap, beta=beta, len_band=len_band)


class SiStemStrKernel(_Kernel):
    kind = L.SI_STEM_STR

    def __init__(self, loop_gap=0.2, stack=1.3, covar=0.8, gap=0.8, match=1.0, mismatch=0.8, len_band=10, device=0):
        super().__init__(device, loop_gap=loop_gap, stack=stack, covar=covar, gap=gap, match=match, mismatch=mismatch,
                         len_band=len_band)


class SuStemStrKernel(_Kernel):
    kind = L.SU_STEM_STR

    def __init__(self, alpha=0.2, beta=0.3, loop_gap=0.2, gap=0.8, len_band=10, device=0):
        super().__init__(device, alpha=alpha, beta=beta, loop_gap=loop_gap, gap=gap, len_band=len_band)


class LSuStemKernel(_Kernel):
    kind = L.LSU_STEM

    def __init__(self, loop_gap=0.2, beta=0.3, len_band=10, device=0):
        super().__init__(device, loop_gap=loop_gap, beta=beta, len_band=len_band)


class LSuStrKernel(_Kernel):
    kind = L.LSU_STR

    def __init__(self, gap=0.8, alpha=0.2, device=0):
        super().__init__(device, gap=gap, alpha=alpha)


class LSuStemStrKernel(_Kernel):
    kind = L.LSU_STEM_STR

    def __init__(self, alpha=0.2, beta=0.3, loop_gap=0.2, gap=0.8, len_band=10, device=0):
        super().__init__(device, alpha=alpha, beta=beta, loop_gap=loop_gap, gap=gap, len_band=len_band)


class StringKernel(_Kernel):
    """Lite string kernel: StringKernel(gap, alpha) or StringKernel(gap, match, mismatch)."""

    def __init__(self, gap=0.8, alpha=None, match=None, mismatch=None, device=0):
        if alpha is not None:
            self.kind = L.STR_SUBST
            super().__init__(device, gap=gap, alpha=alpha)
        else:
            self.kind = L.STR_SIMPLE
            super().__init__(device, gap=gap, match=1.0 if match is None else match,
                             mismatch=0.8 if mismatch is None else mismatch)


class NaiveStringKernel(_Kernel):
    """string_kernel/ binary: the gap is parsed as a float and widened (string_kernel/main.cpp:26,40,93)."""
    kind = L.STR_NAIVE

    def __init__(self, gap=1.0, device=0):
        super().__init__(device, gap=float(np.float32(gap)))


# ---------------------------------------------------------------- KernelMatrix (kernel_matrix.h)
class KernelMatrix:
    def __init__(self):
        self.matrix = np.zeros((0, 0))
        self.labels = []
        self.self_ = np.zeros(0)

    def calculate(self, train, kernel, normalize=False, n_th=1):
        """calculate(train, kernel, normalize, n_th): train = list of (label, MData).  n_th is accepted for
        signature compatibility; the work is scheduled over the GPU's CTAs instead of host threads."""
        dset = kernel.ctx.upload([d for _, d in train])
        self.matrix = kernel.ctx.gram(dset, normalize)
        self.labels = [lab for lab, _ in train]
        return self

    def calculate_test(self, test, train, kernel, norm_test=False, normalize=False, n_th=1, sv_index=None):
        """calculate(test, train, kernel, norm_test, normalize, n_th) (+ the sv_index of the row variant)."""
        dtest = kernel.ctx.upload([d for _, d in test])
        dtrain = kernel.ctx.upload([d for _, d in train])
        self.matrix, selfv = kernel.ctx.cross(dtest, dtrain, sv_index, normalize, want_self=norm_test or normalize)
        self.self_ = selfv
        self.labels = [lab for lab, _ in test]
        return self

    @staticmethod
    def diagonal(train, kernel, sv_index=None, n_th=1):
        dtrain = kernel.ctx.upload([d for _, d in train])
        return kernel.ctx.diag(dtrain, sv_index)

    def print(self, out):
        """KernelMatrix::print (kernel_matrix.cpp:756-770): `label 0:<row> 1:v 2:v ... ` per line,
        values with the C++ stream default of 6 significant digits."""
        out.write(format_matrix(self.matrix, self.labels))


def format_rows(m, labels, first_cnt=1, n_threads=0):
    """Native (threaded) writer of kernel-matrix rows in the reference's text format: bytes of
    "<label> 0:<cnt> 1:<v> ... \\n" lines (KernelMatrix::print, kernel_matrix.cpp:756-770; Output::kernel_output,
    framework.cpp:190-204).  Byte-identical to format_matrix() below, which stays as the independent check."""
    import ctypes as C
    m = np.ascontiguousarray(m, dtype=np.float64)
    if m.ndim == 1:
        m = m[None, :]
    n_rows, n_cols = m.shape
    if n_rows == 0:
        return b""
    lab = (C.c_char_p * n_rows)(*[str(x).encode() for x in labels])
    cap = n_rows * (64 + 26 * n_cols)
    buf = C.create_string_buffer(cap)
    n = L.lib().stemk_format_rows(m.ctypes.data, n_rows, n_cols, n_cols, lab, first_cnt, n_threads, buf, cap)
    if n > cap:
        buf = C.create_string_buffer(n)
        n = L.lib().stemk_format_rows(m.ctypes.data, n_rows, n_cols, n_cols, lab, first_cnt, n_threads, buf, n)
    return buf.raw[:n]


def format_values(v):
    """One "%g" value per line (the norm file, Output::norm_output, framework.cpp:218-228)."""
    import ctypes as C
    v = np.ascontiguousarray(v, dtype=np.float64)
    cap = 32 * max(1, len(v))
    buf = C.create_string_buffer(cap)
    n = L.lib().stemk_format_values(v.ctypes.data, len(v), buf, cap)
    return buf.raw[:n]


def stream_predict(ctx, train, test_batches, out, norm_out=None, sv_index=None, normalize=False, train_diag=None):
    """Streaming mirror of App::predict + Output (common/framework.h:167-306, framework.cpp:141-234): for every batch
    (labels, uploaded test set) of `test_batches` the rows k(test_i, train_j) are computed on the device
    (stemk_cross), normalised like framework.h:279-283 (`vec[j] /= sqrt(diag[j]*self)`), written to the binary file
    object `out` in the reference's text format and, if given, the self terms to `norm_out`.  The text of batch b is
    formatted by the native threaded writer while the device computes batch b+1.  Yields (labels, rows, self) per
    batch so that a caller can feed the unchanged CPU svm_predict.  Returns nothing else: the matrix is never held
    in memory as a whole."""
    import threading
    norm = normalize or norm_out is not None
    diag = None
    if normalize:
        diag = np.asarray(train_diag) if train_diag is not None else ctx.diag(train, sv_index)
    cnt = 1
    pending = None

    def flush(job):
        labels, rows, selfv, first = job
        out.write(format_rows(rows, labels, first))
        if norm_out is not None:
            norm_out.write(format_values(selfv))

    for labels, test in test_batches:
        rows, selfv = ctx.cross(test, train, sv_index, normalize=False, want_self=norm)
        if normalize:
            rows = rows / np.sqrt(diag[None, :] * selfv[:, None])
        if pending is not None:
            pending[0].join()
            yield pending[1]
        job = (list(labels), rows, selfv, cnt)
        th = threading.Thread(target=flush, args=(job,))
        th.start()
        pending = (th, (job[0], rows, selfv))
        cnt += len(job[0])
    if pending is not None:
        pending[0].join()
        yield pending[1]


def format_matrix(m, labels):
    lines = []
    for i in range(m.shape[0]):
        parts = [f"{labels[i]} 0:{i + 1} "]
        parts.extend(f"{j + 1}:{_g6(m[i, j])} " for j in range(m.shape[1]))
        lines.append("".join(parts) + "\n")
    return "".join(lines)


def _g6(v):
    if v != v:   # glibc (and with it libstdc++'s operator<<) spells a NaN with the sign bit set "-nan"; x86 0/0 is one
        return "-nan" if np.signbit(v) else "nan"
    return "%g" % v
