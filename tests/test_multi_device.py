"""Several devices behind the C ABI (reference: common/kernel_matrix.cpp:186-262, 495-526, 560-571):
stemk_set_clone / stemk_upload_multi / stemk_gram_multi for one process driving several devices, and
stemk_set_export / stemk_set_import for one process per device.  The C program tests/c/gram_multi_test.c is the
"plain C caller" of include/stemk.h; the Python tests go through the ctypes mirror."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stem_kernel_b200 import _lib as L  # noqa: E402
from stem_kernel_b200 import api, hostlib, synth  # noqa: E402


def _records(n3=10, n1=6):
    return synth.make_config(3, n3, offset=300) + synth.make_config(1, n1, offset=400)


def dump_desc(flat, path):
    """The arrays of a stemk_seqset_desc, in declaration order, behind seven uint64 counts."""
    d = flat.desc()
    ns = d.n_seqs
    u32 = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(max(n, 1),))[:n].copy()
    f32 = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(max(n, 1),))[:n].copy()
    u8 = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(n, 1),))[:n].copy()
    node_off = u32(d.node_off, ns + 1)
    nn = int(node_off[-1])
    edge_off = u32(d.edge_off, nn + 1)
    ne = int(edge_off[-1])
    bpf_off = u32(d.bpf_off, nn + 1)
    nb = int(bpf_off[-1])
    root_off = u32(d.root_off, ns + 1)
    nr = int(root_off[-1])
    col_off = u32(d.col_off, ns + 1)
    nc = int(col_off[-1])
    weight_off = u32(d.weight_off, ns + 1)
    nw = int(weight_off[-1])
    parts = [np.array([ns, nn, ne, nb, nr, nc, nw], dtype=np.uint64),
             node_off, u32(d.node_first, nn), u32(d.node_last, nn), f32(d.node_weight, nn),
             edge_off, u32(d.edge_to, ne), u32(d.edge_gaps, ne), f32(d.edge_weight, ne),
             bpf_off, u8(d.bpf_a, nb), u8(d.bpf_b, nb), f32(d.bpf_freq, nb),
             root_off, u32(d.root, nr), col_off, f32(d.profile, 5 * nc), f32(d.n_rows, ns),
             weight_off, f32(d.col_weight, nw), u8(d.text, nc)]
    with open(path, "wb") as f:
        for a in parts:
            f.write(np.ascontiguousarray(a).tobytes())


def test_dump_desc_layout(tmp_path):
    """Host only: the dump the C test reads has the size its seven counts announce."""
    flat = hostlib.SeqSet([hostlib.MData.from_record(r) for r in _records(3, 2)])
    p = tmp_path / "desc.bin"
    dump_desc(flat, str(p))
    raw = p.read_bytes()
    ns, nn, ne, nb, nr, nc, nw = np.frombuffer(raw[:56], dtype=np.uint64).astype(int)
    want = 56 + 4 * (ns + 1) + 12 * nn + 4 * (nn + 1) + 12 * ne + 4 * (nn + 1) + 6 * nb + 4 * (ns + 1) + 4 * nr \
        + 4 * (ns + 1) + 20 * nc + 4 * ns + 4 * (ns + 1) + 4 * nw + nc
    assert len(raw) == want and ns == 5


@pytest.mark.gpu
def test_c_caller_gram_multi(tmp_path):
    """A C program linked against libstemk_b200.so only: stemk_upload_multi + stemk_gram_multi over every device of
    the box (2 when there are at least 2) give stemk_gram's matrix bit for bit; so does an exported + imported set."""
    import torch
    flat = hostlib.SeqSet([hostlib.MData.from_record(r) for r in _records()])
    desc = tmp_path / "desc.bin"
    dump_desc(flat, str(desc))
    exe = tmp_path / "gram_multi_test"
    libdir = os.path.join(ROOT, "stem_kernel_b200", "csrc")
    cuda_lib = "/usr/local/cuda/lib64"
    cmd = ["gcc", "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "gram_multi_test.c"),
           "-o", str(exe), "-L", libdir, "-l:libstemk_b200.so", "-L", cuda_lib, "-lcudart",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    n_dev = min(2, torch.cuda.device_count())
    r = subprocess.run([str(exe), str(desc), str(n_dev)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GRAM_MULTI OK" in r.stdout, r.stdout + r.stderr
    assert f"devices {n_dev}" in r.stdout


@pytest.mark.gpu
def test_clone_export_import_and_multi_python():
    """The ctypes mirror: clone onto a second context, export/import, gram_multi over the available devices."""
    import torch
    md = [hostlib.MData.from_record(r) for r in _records()]
    p = L.make_params(L.SU_STEM)
    n_dev = min(2, torch.cuda.device_count())
    ctxs = [api.Context(p, device=d) for d in range(n_dev)]
    sets = api.upload_multi(ctxs, md)
    want = ctxs[0].gram(sets[0], normalize=True)
    got = api.gram_multi(ctxs, sets, normalize=True)
    assert np.array_equal(got, want)
    # a clone on the SAME device through a second context (a peer copy whose two ends coincide)
    ctx2 = api.Context(p, device=0)
    cl = sets[0].clone_to(ctx2)
    assert np.array_equal(ctx2.gram(cl, normalize=True), want)
    # export -> import
    buf = torch.empty(sets[0].export_bytes(), dtype=torch.uint8, device="cuda:0")
    sets[0].export_to(buf.data_ptr())
    imp = api.DeviceSet.import_from(ctx2, buf.data_ptr(), buf.numel())
    assert len(imp) == len(sets[0])
    assert np.array_equal(ctx2.gram(imp, normalize=True), want)
    v0, e0, l0 = sets[0].stats()
    v1, e1, l1 = imp.stats()
    assert np.array_equal(v0, v1) and np.array_equal(e0, e1) and np.array_equal(l0, l1)
    # a set compiled under another loop gap is refused
    ctx3 = api.Context(L.make_params(L.SU_STEM, loop_gap=0.5), device=0)
    with pytest.raises(api.StemkError):
        api.DeviceSet.import_from(ctx3, buf.data_ptr(), buf.numel())
    with pytest.raises(api.StemkError):
        sets[0].clone_to(ctx3)
