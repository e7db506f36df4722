"""ctypes binding of libstemk_host.so: the host front end (rows + base-pair lists -> MData) and the
flattening adapter (list of MData -> stemk_seqset_desc).  Pure host code, no CUDA."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(_HERE, "host", "libstemk_host.so")


class SeqSetDesc(C.Structure):
    """stemk_seqset_desc of include/stemk.h."""
    _fields_ = [("n_seqs", C.c_uint32)] + [(n, C.c_void_p) for n in (
        "node_off", "node_first", "node_last", "node_weight", "edge_off", "edge_to", "edge_gaps", "edge_weight",
        "bpf_off", "bpf_a", "bpf_b", "bpf_freq", "root_off", "root", "col_off", "profile", "n_rows", "weight_off",
        "col_weight", "text")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_SO):
            raise ImportError(f"{HOST_SO} is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(HOST_SO)
        vp, u32p = C.c_void_p, C.POINTER(C.c_uint32)
        L.stemk_host_last_error.restype = C.c_char_p
        L.stemk_host_mdata_build.restype = vp
        L.stemk_host_mdata_build.argtypes = [C.c_int, C.POINTER(C.c_char_p), vp, vp, vp, vp, C.c_float]
        L.stemk_host_mdata_seqonly.restype = vp
        L.stemk_host_mdata_seqonly.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.stemk_host_mdata_build_many.restype = C.c_int
        L.stemk_host_mdata_build_many.argtypes = [C.c_int, C.POINTER(C.c_char_p), vp, vp, vp, vp, C.c_float, C.c_int,
                                                  C.POINTER(vp)]
        L.stemk_host_mdata_free.argtypes = [vp]
        L.stemk_host_mdata_sizes.argtypes = [vp, u32p]
        L.stemk_host_mdata_export.argtypes = [vp] * 17
        L.stemk_host_mdata_from_arrays.restype = vp
        L.stemk_host_mdata_from_arrays.argtypes = [C.c_uint32] + [vp] * 11 + [C.c_uint32, vp, C.c_uint32, vp,
                                                                             C.c_float, C.c_uint32, vp, C.c_char_p]
        L.stemk_host_set_new.restype = vp
        L.stemk_host_set_free.argtypes = [vp]
        L.stemk_host_set_add.argtypes = [vp, vp]
        L.stemk_host_set_size.restype = C.c_uint32
        L.stemk_host_set_size.argtypes = [vp]
        L.stemk_host_set_desc.argtypes = [vp, C.POINTER(SeqSetDesc)]
        _lib = L
    return _lib


def _rows(rows):
    return (C.c_char_p * len(rows))(*[r.encode() if isinstance(r, str) else r for r in rows])


class MData:
    """One structure-annotated record (mirror of the reference's MData, stem_kernel_lite/data.h:26-53)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError(lib().stemk_host_last_error().decode())
        self.h = handle

    @classmethod
    def build(cls, rows, bp_rows, th=0.01):
        """rows: aligned strings; bp_rows: per row (bi, bj, bp), 1-based over the ungapped row."""
        off = np.zeros(len(rows) + 1, dtype=np.uint32)
        for k, (bi, _, _) in enumerate(bp_rows):
            off[k + 1] = off[k] + len(bi)
        bi = np.ascontiguousarray(np.concatenate([np.asarray(b[0]) for b in bp_rows]), dtype=np.uint32)
        bj = np.ascontiguousarray(np.concatenate([np.asarray(b[1]) for b in bp_rows]), dtype=np.uint32)
        bp = np.ascontiguousarray(np.concatenate([np.asarray(b[2]) for b in bp_rows]), dtype=np.float64)
        return cls(lib().stemk_host_mdata_build(len(rows), _rows(rows), off.ctypes.data, bi.ctypes.data,
                                                bj.ctypes.data, bp.ctypes.data, C.c_float(th)))

    @classmethod
    def seq_only(cls, rows):
        return cls(lib().stemk_host_mdata_seqonly(len(rows), _rows(rows)))

    @classmethod
    def from_record(cls, rec, th=0.01, structure=True):
        return cls.build(rec["rows"], rec["bp"], th) if structure else cls.seq_only(rec["rows"])

    @classmethod
    def from_arrays(cls, f, text=""):
        """f: dict as returned by export() (tests feed hand-made DAGs through this)."""
        a = lambda k, dt: np.ascontiguousarray(f[k], dtype=dt)
        u32, f32, u8 = np.uint32, np.float32, np.uint8
        keep = [a("first", u32), a("last", u32), a("weight", f32), a("edge_off", u32), a("edge_to", u32),
                a("edge_gaps", u32), a("edge_w", f32), a("bpf_off", u32), a("bpf_a", u8), a("bpf_b", u8),
                a("bpf_f", f32)]
        root, prof, sw = a("root", u32), a("profile", f32), a("seq_weight", f32)
        h = lib().stemk_host_mdata_from_arrays(len(keep[0]), *[k.ctypes.data for k in keep], len(root),
                                               root.ctypes.data, prof.shape[0], prof.ctypes.data,
                                               C.c_float(f["n_seqs"]), len(sw), sw.ctypes.data, text.encode())
        return cls(h)

    def __del__(self):
        try:
            if self.h:
                lib().stemk_host_mdata_free(self.h)
                self.h = None
        except Exception:
            pass

    def sizes(self):
        s = (C.c_uint32 * 6)()
        lib().stemk_host_mdata_sizes(self.h, s)
        return dict(zip(("n_nodes", "n_edges", "n_bpf", "n_roots", "length", "n_weights"), list(s)))

    def export(self):
        s = self.sizes()
        nn, ne, nb, nr, L, nw = (s[k] for k in ("n_nodes", "n_edges", "n_bpf", "n_roots", "length", "n_weights"))
        u32, f32, u8 = np.uint32, np.float32, np.uint8
        d = dict(first=np.zeros(nn, u32), last=np.zeros(nn, u32), weight=np.zeros(nn, f32),
                 edge_off=np.zeros(nn + 1, u32), edge_to=np.zeros(ne, u32), edge_gaps=np.zeros(ne, u32),
                 edge_w=np.zeros(ne, f32), bpf_off=np.zeros(nn + 1, u32), bpf_a=np.zeros(nb, u8),
                 bpf_b=np.zeros(nb, u8), bpf_f=np.zeros(nb, f32), root=np.zeros(nr, u32), max_pa=np.zeros(nn, u32),
                 profile=np.zeros((L, 5), f32), n_seqs=np.zeros(1, f32), seq_weight=np.zeros(nw, f32))
        order = ["first", "last", "weight", "edge_off", "edge_to", "edge_gaps", "edge_w", "bpf_off", "bpf_a", "bpf_b",
                 "bpf_f", "root", "max_pa", "profile", "n_seqs", "seq_weight"]
        lib().stemk_host_mdata_export(self.h, *[d[k].ctypes.data for k in order])
        d["n_seqs"] = float(d["n_seqs"][0])
        return d


def build_many(records, th=0.01, n_threads=None):
    """Single-row records -> list[MData], built by a C++ thread pool."""
    n = len(records)
    if n == 0:
        return []
    if any(len(r["rows"]) != 1 for r in records):
        return [MData.from_record(r, th) for r in records]
    n_threads = n_threads or min(32, os.cpu_count() or 1)
    off = np.zeros(n + 1, dtype=np.uint64)
    for k, r in enumerate(records):
        off[k + 1] = off[k] + len(r["bp"][0][0])
    bi = np.ascontiguousarray(np.concatenate([np.asarray(r["bp"][0][0]) for r in records]), dtype=np.uint32)
    bj = np.ascontiguousarray(np.concatenate([np.asarray(r["bp"][0][1]) for r in records]), dtype=np.uint32)
    bp = np.ascontiguousarray(np.concatenate([np.asarray(r["bp"][0][2]) for r in records]), dtype=np.float64)
    out = (C.c_void_p * n)()
    rc = lib().stemk_host_mdata_build_many(n, _rows([r["rows"][0] for r in records]), off.ctypes.data, bi.ctypes.data,
                                           bj.ctypes.data, bp.ctypes.data, C.c_float(th), n_threads, out)
    if rc != 0:
        raise ValueError(lib().stemk_host_last_error().decode())
    return [MData(h) for h in out]


class SeqSet:
    """Flattened list of MData: owns the arrays a stemk_seqset_desc points into."""

    def __init__(self, mdatas=()):
        self.h = lib().stemk_host_set_new()
        self._keep = []
        for m in mdatas:
            self.add(m)

    def add(self, m):
        lib().stemk_host_set_add(self.h, m.h)

    def __len__(self):
        return lib().stemk_host_set_size(self.h)

    def desc(self):
        d = SeqSetDesc()
        lib().stemk_host_set_desc(self.h, C.byref(d))
        d._owner = self  # the descriptor points into this set's arrays: keep them alive with it
        return d

    def __del__(self):
        try:
            if self.h:
                lib().stemk_host_set_free(self.h)
                self.h = None
        except Exception:
            pass
