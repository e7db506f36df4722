"""Base-pair probabilities on the device: ctypes mirror of stemk_fold_bpp / stemk_fold_fetch.

Stands where the reference's front end calls ViennaRNA once per sequence under a mutex (common/bpmatrix.cpp:141-177):
a batch of sequences in, the sparse per-sequence lists (i, j, p) of BPMatrix entries out -- what hostlib.MData.from_record
(Profiler + DAGBuilder, host/frontend.cpp) consumes.  The energy model is an input (FoldModel = stemk_fold_model of
include/stemk.h); parity with ViennaRNA is unpinned, the checker is oracle/stemk_fold_oracle.c."""
import ctypes as C

import numpy as np

from . import _lib as L
from .api import Context, StemkError


class FoldModel(C.Structure):
    """stemk_fold_model of include/stemk.h (energies in kcal/mol)."""
    _fields_ = [("temperature", C.c_double), ("pf_scale", C.c_double), ("no_gu", C.c_int32), ("pad_", C.c_int32),
                ("stack", C.c_double * 8 * 8), ("hairpin", C.c_double * 31), ("bulge", C.c_double * 31),
                ("interior", C.c_double * 31), ("lxc", C.c_double), ("mismatch_h", C.c_double * 5 * 5 * 8),
                ("mismatch_i", C.c_double * 5 * 5 * 8), ("dangle5", C.c_double * 5 * 8), ("dangle3", C.c_double * 5 * 8),
                ("ninio", C.c_double), ("max_ninio", C.c_double), ("terminal_au", C.c_double), ("ml_closing", C.c_double),
                ("ml_intern", C.c_double * 8), ("ml_base", C.c_double)]


def default_model():
    """stemk_fold_model_default: the stand-in parameter set (not a published parameter file)."""
    m = FoldModel()
    L.lib().stemk_fold_model_default(C.byref(m))
    return m


class FoldResult:
    """pairs[k] = (i, j, p) arrays of sequence k (1-based, i < j, ascending (i, j)); unpaired[k] = per-position
    max(0, 1 - sum_j P); ensemble[k] = -kT ln Z; dense[k] = (L+1) x (L+1) table when asked for."""

    def __init__(self, pairs, unpaired, ensemble, dense, kernel_ms=0.0):
        self.pairs, self.unpaired, self.ensemble, self.dense, self.kernel_ms = pairs, unpaired, ensemble, dense, kernel_ms


class Folder:
    """A context used for the front end only (the kernel parameters play no role in it)."""

    def __init__(self, ctx=None, device=0):
        self._own = ctx is None
        self.ctx = ctx if ctx is not None else Context(L.make_params(L.SU_STEM), device=device)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self._own and self.ctx is not None:
            self.ctx.close()
        self.ctx = None

    def bpp(self, seqs, model=None, cutoff=1e-5, dense=False):
        model = model if model is not None else default_model()
        # anything that is not ASCII cannot be a base: keep one byte per character
        text = "".join(seqs).encode("ascii", errors="replace")
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        total = C.c_uint64(0)
        ens = np.zeros(len(seqs))
        dn = np.zeros(int(((lens + 1) ** 2).sum()), dtype=np.float64) if dense else None
        lib = L.lib()
        self.ctx._check(lib.stemk_fold_bpp(self.ctx.h, C.byref(model), len(seqs), off.ctypes.data, text, float(cutoff), C.byref(total),
                                           ens.ctypes.data, dn.ctypes.data if dense else None))
        n = int(total.value)
        poff = np.zeros(len(seqs) + 1, dtype=np.uint64)
        bi, bj, bp = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32), np.zeros(n)
        unp = np.zeros(int(off[-1]))
        self.ctx._check(lib.stemk_fold_fetch(self.ctx.h, poff.ctypes.data, bi.ctypes.data, bj.ctypes.data, bp.ctypes.data, unp.ctypes.data))
        pairs, unpaired, dl = [], [], []
        dpos = 0
        for k in range(len(seqs)):
            a, b = int(poff[k]), int(poff[k + 1])
            pairs.append((bi[a:b].astype(np.int64), bj[a:b].astype(np.int64), bp[a:b]))
            unpaired.append(unp[int(off[k]):int(off[k + 1])])
            if dense:
                w = int(lens[k]) + 1
                dl.append(dn[dpos:dpos + w * w].reshape(w, w))
                dpos += w * w
        return FoldResult(pairs, unpaired, ens, dl if dense else None, float(lib.stemk_fold_last_ms(self.ctx.h)))


def build_mdata(seqs, th=0.01, model=None, cutoff=0.0, folder=None, n_threads=None):
    """The reference's Data(const IS&, float th, ...) constructor for a batch of single sequences (data.cpp:324-345) with
    the device in the place of ViennaRNA: base-pair probabilities of all sequences in one stemk_fold_bpp call, then
    Profiler + DAGBuilder on the host threads (hostlib.build_many).  cutoff: pairs below it are not handed to the front
    end (0 keeps every pair, like the dense BPMatrix the reference reads; any value below th gives the same DAG)."""
    from . import hostlib
    own = folder is None
    folder = folder if folder is not None else Folder()
    try:
        res = folder.bpp(seqs, model, cutoff=cutoff)
    finally:
        if own:
            folder.close()
    recs = [dict(rows=[s], bp=[res.pairs[k]], label=1) for k, s in enumerate(seqs)]
    return hostlib.build_many(recs, th, n_threads)


__all__ = ["FoldModel", "FoldResult", "Folder", "build_mdata", "default_model", "StemkError"]
