"""Host-only context (device = STEMK_DEVICE_NONE): record compilation (compile_set.cpp) and the SURVEY 8(d) work
model run without a GPU; anything that would compute a kernel value refuses."""
import numpy as np
import pytest

from conftest import TH
from oracle import oraclebind as O
from stem_kernel_b200 import _lib as L
from stem_kernel_b200 import api, hostlib, synth


def test_set_stats_and_cost_model_equal_oracle(golden):
    recs = synth.make_config(3, 5, offset=70)
    md = golden["md"] + [hostlib.MData.from_record(r, TH) for r in recs]
    flat = hostlib.SeqSet(md)
    d = flat.desc()
    n = len(md)
    xi, yi = np.divmod(np.arange(n * n), n)
    for kind, band in ((L.SU_STEM, 10), (L.SU_STEM, 0), (L.SU_STEM_STR, 3), (L.STR_SUBST, 0), (L.STR_NAIVE, 0)):
        p = L.make_params(kind, len_band=band)
        ctx = api.Context(p, device=-1)
        ds = ctx.upload(flat)
        v, e, length = ds.stats()
        assert [int(x) for x in v] == [m.sizes()["n_nodes"] for m in md]
        assert [int(x) for x in e] == [m.sizes()["n_edges"] for m in md]
        assert [int(x) for x in length] == [m.sizes()["length"] for m in md]
        cells, flops = ctx.pair_cost(ds, ds, xi, yi)
        for k in range(0, n * n, 7):
            c, f = O.pair_cost(O.Params.from_buffer_copy(p), d, int(xi[k]), d, int(yi[k]))
            assert (cells[k], flops[k]) == (c, f), (kind, band, k)
        with pytest.raises(api.StemkError, match="no CPU path"):
            ctx.gram(ds)
        with pytest.raises(api.StemkError, match="no CPU path"):
            ctx.pairs(ds, ds, [0], [0])


def test_upload_rejects_malformed_records():
    f = dict(first=[0, 3], last=[0, 9], weight=[1.0, 1.0], edge_off=[0, 1, 1], edge_to=[1], edge_gaps=[0],
             edge_w=[1.0], bpf_off=[0, 0, 0], bpf_a=[], bpf_b=[], bpf_f=[], root=[1],
             profile=np.zeros((10, 5), np.float32), n_seqs=1.0, seq_weight=[])
    bad = hostlib.MData.from_arrays(f, "a" * 10)       # node 0 lists node 1 as a child: parents before children
    ctx = api.Context(L.make_params(L.SU_STEM), device=-1)
    with pytest.raises(api.StemkError, match="children before parents"):
        ctx.upload([bad])


@pytest.mark.parametrize("case,msg", [("last_before_first", "first <= last"), ("last_outside", "first <= last"),
                                      ("huge_gaps", "gap count"), ("no_rows", "n_rows")])
def test_upload_validates_the_descriptor(case, msg):
    """Hostile descriptors are refused with STEMK_ERR_ARG instead of sizing arrays from wrapped-around numbers
    (ADVICE round 1): node_last < node_first, positions outside the sequence, absurd gap counts, n_rows = 0."""
    f = dict(first=[2, 1], last=[2, 8], weight=[1.0, 1.0], edge_off=[0, 0, 1], edge_to=[0], edge_gaps=[6],
             edge_w=[1.0], bpf_off=[0, 0, 1], bpf_a=[0], bpf_b=[3], bpf_f=[1.0], root=[1],
             profile=np.zeros((10, 5), np.float32), n_seqs=1.0, seq_weight=[])
    ctx = api.Context(L.make_params(L.SU_STEM), device=-1)
    ctx.upload([hostlib.MData.from_arrays(f, "a" * 10)])          # the well-formed record is accepted
    if case == "last_before_first":
        f["last"] = [2, 0]
    elif case == "last_outside":
        f["last"] = [2, 10]
    elif case == "huge_gaps":
        f["edge_gaps"] = [4000000000]
    else:
        f["n_seqs"] = 0.0
    with pytest.raises(api.StemkError, match=msg):
        ctx.upload([hostlib.MData.from_arrays(f, "a" * 10)])


def test_fold_has_no_cpu_path_and_checks_its_arguments():
    """stemk_fold_bpp computes on a CUDA device only; a host-only context refuses, null arguments are STEMK_ERR_ARG."""
    import ctypes as C
    from stem_kernel_b200 import fold
    ctx = api.Context(L.make_params(L.SU_STEM), device=-1)
    with pytest.raises(api.StemkError, match="no CPU path"):
        fold.Folder(ctx).bpp(["gggaaaccc"])
    lib = L.lib()
    total = C.c_uint64(0)
    assert lib.stemk_fold_bpp(ctx.h, None, 0, None, None, 0.0, C.byref(total), None, None) == L.ERR_ARG
    assert lib.stemk_fold_fetch(ctx.h, None, None, None, None, None) == L.ERR_ARG        # nothing to fetch yet
    m = fold.default_model()
    assert m.temperature == 37.0 and m.stack[1][2] == -3.3 and m.hairpin[3] == 5.7 and m.ml_closing == 3.4
    # the stand-in parameter set of the product and the checker's restatement of it are the same numbers
    mo = O.fold_model_default()
    assert bytes(m) == bytes(mo)
