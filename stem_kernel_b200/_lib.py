"""Loader of libstemk_b200.so (CUDA kernels + the C ABI of include/stemk.h).

Fails loudly: there is no CPU implementation behind this package.  If the shared object is missing
the import error says how to build it; if there is no CUDA device, Context() raises."""
import ctypes as C
import os

from .hostlib import SeqSetDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_SO = os.environ.get("STEMK_SO") or os.path.join(_HERE, "csrc", "libstemk_b200.so")   # STEMK_SO: tuning builds

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_STATE = 0, -1, -2, -3, -4
OPT_FORCE_GENERAL, OPT_TIMING, OPT_FORCE_UNSTAGED = 1, 2, 3

# stemk_kind
SI_STEM, SU_STEM, SI_STEM_STR, SU_STEM_STR, LSU_STEM, LSU_STR, LSU_STEM_STR, STR_SUBST, STR_SIMPLE, STR_NAIVE = range(10)
KIND_NAMES = ["SiStemKernel", "SuStemKernel", "SiStemStrKernel", "SuStemStrKernel", "LSuStemKernel", "LSuStrKernel",
              "LSuStemStrKernel", "StringKernel(subst)", "StringKernel(simple)", "string_kernel(naive)"]


class Params(C.Structure):
    """stemk_params."""
    _fields_ = [("kind", C.c_int32), ("len_band", C.c_uint32), ("loop_gap", C.c_double), ("beta", C.c_double),
                ("stack", C.c_double), ("covar", C.c_double), ("gap", C.c_double), ("alpha", C.c_double),
                ("match", C.c_double), ("mismatch", C.c_double)]


def make_params(kind, loop_gap=0.2, beta=0.3, stack=1.3, covar=0.8, gap=0.8, alpha=0.2, match=1.0, mismatch=0.8,
                len_band=10):
    """Defaults are the reference's command-line defaults (stem_kernel_lite/main.cpp:103-149)."""
    return Params(kind, len_band, loop_gap, beta, stack, covar, gap, alpha, match, mismatch)


EXPORTS = ["stemk_version", "stemk_device_count", "stemk_create", "stemk_set_option", "stemk_destroy", "stemk_last_error", "stemk_upload",
           "stemk_set_free", "stemk_set_size", "stemk_set_stats", "stemk_set_device_bytes", "stemk_gram", "stemk_cross", "stemk_diag", "stemk_pairs",
           "stemk_pairs_device", "stemk_assemble_device", "stemk_pair_cost", "stemk_stats_reset", "stemk_stats_get", "stemk_fp64_peak", "stemk_format_rows", "stemk_format_values", "stemk_bpla_pairs", "stemk_bpla_gradients", "stemk_nstem_pairs", "stemk_nstem_pairs_banded", "stemk_nstem_pairs_windows",
           "stemk_set_clone", "stemk_upload_multi", "stemk_gram_multi", "stemk_set_export_bytes", "stemk_set_export", "stemk_set_import",
           "stemk_fold_model_default", "stemk_fold_bpp", "stemk_fold_fetch", "stemk_fold_last_ms"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(CUDA_SO):
            raise ImportError(f"{CUDA_SO} is not built (no CPU fallback exists): run "
                              "`python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(CUDA_SO)
        vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
        L.stemk_version.restype = C.c_char_p
        L.stemk_device_count.restype = C.c_int
        L.stemk_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.c_int]
        L.stemk_destroy.argtypes = [vp]
        L.stemk_set_option.argtypes = [vp, C.c_int, C.c_int]
        L.stemk_last_error.restype = C.c_char_p
        L.stemk_last_error.argtypes = [vp]
        L.stemk_upload.argtypes = [vp, C.POINTER(SeqSetDesc), C.POINTER(vp)]
        L.stemk_set_free.argtypes = [vp, vp]
        L.stemk_set_size.restype = u32
        L.stemk_set_size.argtypes = [vp]
        L.stemk_set_stats.argtypes = [vp, vp, vp, vp]
        L.stemk_set_device_bytes.restype = C.c_uint64
        L.stemk_set_device_bytes.argtypes = [vp]
        L.stemk_gram.argtypes = [vp, vp, C.c_int, vp]
        L.stemk_cross.argtypes = [vp, vp, vp, vp, u32, C.c_int, vp, vp]
        L.stemk_diag.argtypes = [vp, vp, vp, u32, vp]
        L.stemk_pairs.argtypes = [vp, vp, vp, sz, vp, vp, vp]
        L.stemk_pairs_device.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp]
        L.stemk_assemble_device.argtypes = [vp, sz, vp, vp, vp, u32, C.c_int, vp, vp]
        L.stemk_pair_cost.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp]
        L.stemk_stats_reset.argtypes = [vp]
        L.stemk_stats_get.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.stemk_fp64_peak.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
        L.stemk_format_rows.restype = sz
        L.stemk_format_rows.argtypes = [vp, u32, u32, sz, C.POINTER(C.c_char_p), u32, C.c_int, vp, sz]
        L.stemk_bpla_pairs.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
        L.stemk_bpla_gradients.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp]
        L.stemk_nstem_pairs.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
        L.stemk_nstem_pairs_banded.argtypes = [vp, vp, C.c_uint32, vp, vp, sz, vp, vp, vp]
        L.stemk_nstem_pairs_windows.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp, vp, vp]
        L.stemk_set_clone.argtypes = [vp, vp, C.POINTER(vp)]
        L.stemk_upload_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(SeqSetDesc), C.POINTER(vp)]
        L.stemk_gram_multi.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int, vp]
        L.stemk_set_export_bytes.restype = C.c_uint64
        L.stemk_set_export_bytes.argtypes = [vp]
        L.stemk_set_export.argtypes = [vp, vp, vp, vp]
        L.stemk_set_import.argtypes = [vp, vp, C.c_uint64, C.POINTER(vp)]
        L.stemk_format_values.restype = sz
        L.stemk_format_values.argtypes = [vp, sz, vp, sz]
        L.stemk_fold_model_default.argtypes = [vp]
        L.stemk_fold_bpp.argtypes = [vp, vp, u32, vp, C.c_char_p, C.c_double, C.POINTER(C.c_uint64), vp, vp]
        L.stemk_fold_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
        L.stemk_fold_last_ms.restype = C.c_double
        L.stemk_fold_last_ms.argtypes = [vp]
        _lib = L
    return _lib
