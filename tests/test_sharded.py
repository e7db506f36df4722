"""Host logic of the multi-GPU path (stem_kernel_b200/sharded.py): global pair order, strided deal, gather,
un-deal -- on CPU, including a real world_size-2 run over gloo whose arithmetic is the oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from stem_kernel_b200 import sharded


def test_square_pairs_cover_upper_triangle_once():
    keys = np.array([5.0, 1.0, 9.0, 9.0, 0.0, 3.0])
    xi, yi = sharded.square_pairs(keys)
    assert len(xi) == 21 and np.all(xi <= yi)
    assert len({(int(a), int(b)) for a, b in zip(xi, yi)}) == 21
    order = sharded.size_order(keys)
    assert list(order) == [2, 3, 0, 5, 1, 4]                 # biggest first, stable
    assert (int(xi[0]), int(yi[0])) == (2, 2) and (int(xi[-1]), int(yi[-1])) == (4, 4)
    assert list(yi[:3]) == [2, 2, 2] and list(xi[:3]) == [2, 0, 1]        # y-major; partners big first


def test_cross_pairs_roles_and_sv_subset():
    xi, yi = sharded.cross_pairs([1.0, 2.0], [3.0, 1.0, 2.0], cols=[2, 0])
    assert list(yi) == [1, 1, 0, 0]                           # bigger test record first
    assert list(xi) == [0, 2, 0, 2]                           # train record is x (first argument), bigger first


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_deal_undeal_roundtrip(world):
    n_pairs = 1003
    vals = torch.arange(n_pairs, dtype=torch.float64)
    m = sharded.slab(n_pairs, world)
    g = torch.zeros((world, m), dtype=torch.float64)
    seen = np.zeros(n_pairs, dtype=int)
    for r in range(world):
        idx = sharded.deal(n_pairs, r, world)
        seen[idx] += 1
        g[r, : len(idx)] = vals[idx]
        assert len(idx) in (n_pairs // world, n_pairs // world + 1)
    assert np.all(seen == 1)
    assert torch.equal(sharded.undeal(g, n_pairs), vals)


def test_strided_deal_is_cost_balanced_on_config3_like_sizes():
    rng = np.random.default_rng(3)
    L = rng.integers(150, 301, 1500)
    v = (L * rng.uniform(1.5, 2.5, len(L))).astype(np.int64)
    e = (v * rng.uniform(2.5, 5.0, len(L))).astype(np.int64)
    keys = 2.0 * v * e
    xi, yi = sharded.square_pairs(keys)
    cost = sharded.stem_cost_proxy(v, e, L, xi, yi)
    for world in (2, 4, 8):
        assert sharded.imbalance(cost, world) < 1.005


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo_matches_golden(tmp_path, golden):
    port, out = _free_port(), str(tmp_path / "gram.npy")
    worker = os.path.join(ROOT, "tests", "_sharded_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), out], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    got = np.load(out)
    assert np.array_equal(got, golden["z"]["gram_norm_k3_b10"], equal_nan=True)


def test_world_size_2_gloo_rectangular_matches_the_oracle(tmp_path):
    """ShardedCross (test x train with an sv subset, self terms, train diagonals, one gather) against the oracle's
    one-process cross matrix: bit-identical values, untouched columns outside sv_index turned into the same
    non-finite pattern by the normalisation."""
    from conftest import TH, load_golden_records
    from oracle import oraclebind as O
    from stem_kernel_b200 import _lib as L, hostlib
    port, out = _free_port(), str(tmp_path / "cross.npy")
    worker = os.path.join(ROOT, "tests", "_sharded_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), out, "cross"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    got = np.load(out)
    recs, _ = load_golden_records()
    md = [hostlib.MData.from_record(r, TH) for r in recs]
    ftr, fte = hostlib.SeqSet(md[:11]), hostlib.SeqSet(md[11:])
    p = O.Params.from_buffer_copy(L.make_params(L.SU_STEM_STR))
    cols = np.array([0, 2, 3, 5, 7, 8, 10], dtype=np.uint32)
    want, want_self = O.cross(p, fte.desc(), ftr.desc(), sv_index=cols, normalize=True, init=-3.0)
    assert got.shape == (len(md) - 11, 12)
    assert np.array_equal(got[:, :11], want, equal_nan=True) and np.array_equal(got[:, 11], want_self)
    others = np.setdiff1d(np.arange(11), cols)
    assert not np.isfinite(got[:, others]).any()


def test_world_size_2_gloo_sharded_fold(tmp_path):
    """ShardedFold: a ragged batch of sequences dealt longest-first over two ranks, pair lists gathered on rank 0 in
    the caller's order, equal to the one-process result (the worker compares them; arithmetic = the oracle)."""
    port, out = _free_port(), str(tmp_path / "fold.npy")
    worker = os.path.join(ROOT, "tests", "_sharded_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port), out, "fold"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    counts = np.load(out)
    assert len(counts) == 11 and counts[-2] == 0 and counts.sum() > 0
