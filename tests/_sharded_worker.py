"""One rank of the world_size-2 gloo test of stem_kernel_b200/sharded.py (launched by test_sharded.py).
The arithmetic is the oracle (allowed: this is a test); what is under test is dealing, gather and un-deal."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import TH, load_golden_records  # noqa: E402
from oracle import oraclebind as O  # noqa: E402
from stem_kernel_b200 import _lib as L  # noqa: E402
from stem_kernel_b200 import hostlib, sharded  # noqa: E402


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    recs, _ = load_golden_records()
    md = [hostlib.MData.from_record(r, TH) for r in recs]
    flat = hostlib.SeqSet(md)
    desc = flat.desc()
    p = O.Params.from_buffer_copy(L.make_params(L.SU_STEM_STR))
    sizes = [m.sizes() for m in md]
    keys = [2.0 * s["n_nodes"] * s["n_edges"] + s["length"] ** 2 for s in sizes]

    def compute(xi, yi):
        return torch.from_numpy(O.pairs(p, desc, desc, xi.numpy().view(np.uint32), yi.numpy().view(np.uint32)))

    def assemble(xi, yi, vals, n, normalize):
        m = np.zeros((n, n))
        a, b = xi.numpy().view(np.uint32), yi.numpy().view(np.uint32)
        m[a, b] = vals.numpy()
        m[b, a] = vals.numpy()
        if normalize:
            d = np.diag(m).copy()
            with np.errstate(divide="ignore", invalid="ignore"):
                m = m / np.sqrt(np.outer(d, d))
            np.fill_diagonal(m, 1.0)
        return torch.from_numpy(m)

    if len(sys.argv) > 5 and sys.argv[5] == "fold":
        # front end: base-pair probabilities of a ragged batch, the oracle as arithmetic
        seqs = [r["rows"][0] for r in recs if len(r["rows"]) == 1][:9] + ["", "gggaaaccc"]
        fm = O.fold_model_default()

        def fold_fn(part):
            res = []
            for q in part:
                d = O.fold_bpp(fm, q)[0]
                i, j = np.nonzero(d >= 0.01)
                res.append((i, j, d[i, j]))
            return res

        got = sharded.ShardedFold([len(q) for q in seqs], rank, world).run(seqs, fold_fn)
        assert (got is None) == (rank != 0)
        if rank == 0:
            want = fold_fn(seqs)
            assert len(got) == len(want)
            for g, w in zip(got, want):
                assert all(np.array_equal(a, b) for a, b in zip(g, w))
            np.save(out, np.array([len(g[0]) for g in got]))
        dist.barrier()
        dist.destroy_process_group()
        return
    if len(sys.argv) > 5 and sys.argv[5] == "cross":
        # rectangular: the first 11 golden records are the training set, the rest the test set, 7 sv columns
        n_tr = 11
        ftr, fte = hostlib.SeqSet(md[:n_tr]), hostlib.SeqSet(md[n_tr:])
        descs = {"train": ftr.desc(), "test": fte.desc()}
        cols = np.array([0, 2, 3, 5, 7, 8, 10], dtype=np.uint32)

        def compute_sets(wx, wy, xi, yi):
            return torch.from_numpy(O.pairs(p, descs[wx], descs[wy], xi.numpy().view(np.uint32), yi.numpy().view(np.uint32)))

        sc = sharded.ShardedCross(keys[n_tr:], keys[:n_tr], rank, world, torch.device("cpu"), compute_sets, cols=cols)
        m, selfv = sc.run(normalize=True, init=-3.0)
        assert (m is None) == (rank != 0)
        if rank == 0:
            np.save(out, np.concatenate([m.numpy(), selfv.numpy()[:, None]], axis=1))
        dist.barrier()
        dist.destroy_process_group()
        return
    sg = sharded.ShardedGram(keys, rank, world, torch.device("cpu"), compute, assemble)
    res = sg.run(normalize=True)
    assert (res is None) == (rank != 0)
    if rank == 0:
        np.save(out, res.numpy())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
