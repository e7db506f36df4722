"""Gram matrices over several GPUs of one box: one process per GPU, the pair list dealt out by cost, one gather.

Replaces the reference's MPI path (common/kernel_matrix.cpp:186-262 CalcTrainMatrix with `cnt % size == rank`,
:495-526 Ssend/Recv of each rank's values to rank 0, :560-571 normalisation on the gathered matrix):

  * every rank holds the whole flattened record set in its own HBM (<= 2 GB for 10k records);
  * the n(n+1)/2 pairs (or n_test*n_train for the rectangular matrix) are put into ONE global order -- records
    sorted by a size key, biggest first, pairs y-major (for each y all its partners) -- and rank r owns pairs
    r, r+W, r+2W, ...  Neighbouring pairs of that order cost nearly the same (same y, x adjacent in size), so the
    strided deal is balanced to well under a percent of the stem work model and every rank still runs its own
    expensive pairs first, which is what keeps the tail of its device-side work queue short;
  * the data path has exactly one exchange: a gather of ceil(P/W) doubles per rank to rank 0 (NCCL over NVLink
    on the GPUs, gloo in the CPU tests).  Rank 0 un-deals (a transpose), scatters into the n x n matrix with the
    mirror, and normalises -- all on its device through the C ABI (stemk_assemble_device).

The arithmetic is injected (`compute`, `assemble`): GpuBackend binds them to the C ABI; the CPU tests bind them to
the oracle so that the dealing / gather / un-deal logic is exercised with world_size 2 over gloo.
"""
import numpy as np
import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------------------------- pair orders
def size_order(keys):
    """Record indices, biggest key first (stable), as the single-GPU driver orders them (stemk_api.cu)."""
    keys = np.asarray(keys, dtype=np.float64)
    return np.argsort(-keys, kind="stable").astype(np.uint32)


def square_pairs(keys):
    """All pairs i<=j of a square Gram matrix in the global work order: y-major -- for every record b (biggest
    first) all partners a <= b (biggest first).  The reference evaluates kernel_(x_i, x_j) with i <= j
    (kernel_matrix.cpp:47-50) and the stem kernel is not symmetric, so x = the smaller ORIGINAL index.
    Consecutive pairs share their y record (the stem kernel stages it once per group of pairs)."""
    perm = size_order(keys)
    n = len(perm)
    xs, ys = [], []
    step = max(1, (1 << 24) // max(n, 1))           # rows of the n x n candidate grid per chunk
    for q0 in range(0, n, step):
        b = perm[q0:q0 + step]
        keep = perm[None, :] <= b[:, None]
        ys.append(np.repeat(b, keep.sum(axis=1)))
        xs.append(np.broadcast_to(perm[None, :], keep.shape)[keep])
    if not xs:
        z = np.zeros(0, dtype=np.uint32)
        return z, z
    return np.concatenate(xs).astype(np.uint32), np.concatenate(ys).astype(np.uint32)


def cross_pairs(test_keys, train_keys, cols=None):
    """Pairs of the rectangular matrix: x = train record (FIRST argument, kernel_matrix.cpp:159,168), y = test."""
    cols = np.arange(len(train_keys), dtype=np.uint32) if cols is None else np.asarray(cols, dtype=np.uint32)
    tperm = size_order(test_keys)
    cperm = cols[size_order(np.asarray(train_keys)[cols])]
    yi = np.repeat(tperm, len(cperm)).astype(np.uint32)
    xi = np.tile(cperm, len(tperm)).astype(np.uint32)
    return xi, yi


def deal(n_pairs, rank, world):
    """Global positions owned by `rank`: r, r+W, ...  Every rank gets ceil or floor of P/W pairs."""
    return np.arange(rank, n_pairs, world, dtype=np.int64)


def slab(n_pairs, world):
    """Length of the per-rank gather buffer (ranks with one pair fewer pad with a zero)."""
    return (n_pairs + world - 1) // world


def undeal(gathered, n_pairs):
    """[W, slab] gathered values -> values in global pair order (position k = t*W + r sits at [r, t])."""
    return gathered.t().reshape(-1)[:n_pairs].contiguous()


def stem_cost_proxy(v, e, length, xi, yi, stem=True, string=False):
    """Work model used for balance reports: U_skip = Vx*Ey + Ex*Vy and Lx*Ly string cells."""
    v, e, length = (np.asarray(a, dtype=np.float64) for a in (v, e, length))
    c = np.zeros(len(xi))
    if stem:
        c += v[xi] * e[yi] + e[xi] * v[yi]
    if string:
        c += length[xi] * length[yi]
    return c


def imbalance(cost, world):
    """max over ranks of the dealt cost / mean; 1.0 is perfect."""
    per = np.array([cost[r::world].sum() for r in range(world)])
    return float(per.max() / per.mean()) if per.mean() > 0 else 1.0


# ----------------------------------------------------------------------------------------------- driver
class ShardedGram:
    """Square Gram matrix of one record set over `world` ranks.

    compute(xi, yi) -> 1-D float64 tensor (this rank's values, on `device`)
    assemble(xi_all, yi_all, vals_all, n, normalize) -> n x n tensor (rank 0 only)
    xi/yi are torch uint32-as-int32 tensors on `device` holding record indices.
    """

    def __init__(self, keys, rank, world, device, compute, assemble, group=None):
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.compute, self.assemble = compute, assemble
        self.n = len(keys)
        xi, yi = square_pairs(keys)
        self.n_pairs = len(xi)
        mine = deal(self.n_pairs, rank, world)
        self.n_mine = len(mine)
        as_dev = lambda a: torch.from_numpy(a.view(np.int32).copy()).to(device)
        self.xi_mine, self.yi_mine = as_dev(xi[mine]), as_dev(yi[mine])
        self.slab = slab(self.n_pairs, world)
        self.send = torch.zeros(self.slab, dtype=torch.float64, device=device)
        if rank == 0:
            self.xi_all, self.yi_all = as_dev(xi), as_dev(yi)
            self.recv = torch.zeros((world, self.slab), dtype=torch.float64, device=device)
        self.xi_host, self.yi_host = xi, yi

    def run(self, normalize=False):
        """One Gram matrix.  Returns the n x n tensor on rank 0, None elsewhere."""
        vals = self.compute(self.xi_mine, self.yi_mine)
        self.send[: self.n_mine].copy_(vals)
        if self.world > 1:
            if self.rank == 0:
                dist.gather(self.send, list(self.recv.unbind(0)), dst=0, group=self.group)
            else:
                dist.gather(self.send, None, dst=0, group=self.group)
        else:
            self.recv[0].copy_(self.send)
        if self.rank != 0:
            return None
        return self.assemble(self.xi_all, self.yi_all, undeal(self.recv, self.n_pairs), self.n, normalize)


class ShardedCross:
    """Rectangular test x train matrix over `world` ranks: KernelMatrix::calculate(test, train, ...) and the
    one-row variant behind App::predict (common/kernel_matrix.cpp:635-754, MPI path :264-369; BASELINE config 5:
    5k test x 10k train feeding svm_predict).  Three pair lists are dealt round-robin like the square case --
    the n_test x n_cols kernel rows (x = train record, the FIRST argument, kernel_matrix.cpp:159,168), the n_test
    self terms k(t,t) and, for the normalisation, the n_cols train diagonals (KernelMatrix::diagonal,
    :578-633) -- their values travel to rank 0 in ONE gather, and rank 0 scatters, then divides by
    sqrt(self_i * diag_j) as kernel_matrix.cpp:735-748 / framework.h:279-283 do (columns outside sv_index keep the
    caller's `init` value and a zero diagonal, i.e. they become NaN / inf exactly like the reference's rows).

    compute(which_x, which_y, xi, yi) -> 1-D float64 tensor on `device`; which_* is "train" or "test".
    """

    def __init__(self, test_keys, train_keys, rank, world, device, compute, cols=None, group=None):
        self.rank, self.world, self.device, self.group, self.compute = rank, world, device, group, compute
        self.n_test, self.n_train = len(test_keys), len(train_keys)
        self.cols = np.arange(self.n_train, dtype=np.uint32) if cols is None else np.asarray(cols, dtype=np.uint32)
        xi, yi = cross_pairs(test_keys, train_keys, self.cols)
        tid = size_order(test_keys)
        cid = self.cols[size_order(np.asarray(train_keys)[self.cols])]
        # (which_x, which_y, xi, yi) of the three lists; every list is dealt on its own
        self.lists = [("train", "test", xi, yi), ("test", "test", tid, tid), ("train", "train", cid, cid)]
        self.slabs = [slab(len(l[2]), world) for l in self.lists]
        as_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32).copy()).to(device)
        self.mine = []
        for _, _, a, b in self.lists:
            m = deal(len(a), rank, world)
            self.mine.append((as_dev(a[m]), as_dev(b[m]), len(m)))
        self.send = torch.zeros(sum(self.slabs), dtype=torch.float64, device=device)
        if rank == 0:
            self.recv = torch.zeros((world, sum(self.slabs)), dtype=torch.float64, device=device)
            self.idx = [(torch.from_numpy(a.astype(np.int64)).to(device), torch.from_numpy(b.astype(np.int64)).to(device))
                        for _, _, a, b in self.lists]

    @property
    def n_pairs(self):
        return sum(len(l[2]) for l in self.lists)

    def run(self, normalize=False, init=0.0):
        """Returns (matrix [n_test, n_train], self [n_test]) on rank 0, (None, None) elsewhere."""
        off = 0
        for k, (wx, wy, _, _) in enumerate(self.lists):
            a, b, n = self.mine[k]
            if n and (k < 2 or normalize):
                self.send[off:off + n].copy_(self.compute(wx, wy, a, b))
            off += self.slabs[k]
        if self.world > 1:
            if self.rank == 0:
                dist.gather(self.send, list(self.recv.unbind(0)), dst=0, group=self.group)
            else:
                dist.gather(self.send, None, dst=0, group=self.group)
        else:
            self.recv[0].copy_(self.send)
        if self.rank != 0:
            return None, None
        return self.finish(normalize, init)

    def finish(self, normalize=False, init=0.0):
        """Rank 0, after the gather filled self.recv: un-deal the three lists, scatter, normalise."""
        vals, off = [], 0
        for k, (_, _, a, _) in enumerate(self.lists):
            vals.append(undeal(self.recv[:, off:off + self.slabs[k]], len(a)))
            off += self.slabs[k]
        m = torch.full((self.n_test, self.n_train), float(init), dtype=torch.float64, device=self.device)
        m[self.idx[0][1], self.idx[0][0]] = vals[0]                     # row = test record (y), column = train record (x)
        selfv = torch.zeros(self.n_test, dtype=torch.float64, device=self.device)
        selfv[self.idx[1][0]] = vals[1]
        if normalize:
            diag = torch.zeros(self.n_train, dtype=torch.float64, device=self.device)
            diag[self.idx[2][0]] = vals[2]
            if m.device.type == "cpu":   # torch's CPU sqrt / division are vectorised and not always correctly rounded
                with np.errstate(divide="ignore", invalid="ignore"):
                    m = torch.from_numpy(m.numpy() / np.sqrt(selfv.numpy()[:, None] * diag.numpy()[None, :]))
            else:
                m = m / torch.sqrt(selfv[:, None] * diag[None, :])
        return m, selfv


class GpuCrossBackend:
    """compute(which_x, which_y, xi, yi) of ShardedCross bound to the C ABI (train and test sets uploaded on this
    rank's GPU).  Same stream rule as GpuBackend."""

    def __init__(self, ctx, dtrain, dtest, device):
        self.ctx, self.sets, self.device = ctx, {"train": dtrain, "test": dtest}, device
        self.stream = torch.cuda.Stream(device)

    def compute(self, which_x, which_y, xi, yi):
        s = torch.cuda.current_stream(self.device).cuda_stream
        if not s:
            raise RuntimeError("run under a non-default torch stream (GpuCrossBackend.stream)")
        out = torch.empty(xi.numel(), dtype=torch.float64, device=self.device)
        self.ctx.pairs_device(self.sets[which_x], self.sets[which_y], xi.numel(), xi.data_ptr(), yi.data_ptr(),
                              out.data_ptr(), s)
        return out


class GpuBackend:
    """compute / assemble bound to the C ABI on this rank's GPU, asynchronous on torch's current stream.
    Run ShardedGram.run() under `with torch.cuda.stream(backend.stream)`: the library treats a NULL stream as
    "my own stream", which is not ordered against torch's legacy default stream."""

    def __init__(self, ctx, dset, device):
        self.ctx, self.dset, self.device = ctx, dset, device
        self.matrix = None
        self.stream = torch.cuda.Stream(device)

    def _stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        if not s:
            raise RuntimeError("run under a non-default torch stream (GpuBackend.stream)")
        return s

    def compute(self, xi, yi):
        out = torch.empty(xi.numel(), dtype=torch.float64, device=self.device)
        self.ctx.pairs_device(self.dset, self.dset, xi.numel(), xi.data_ptr(), yi.data_ptr(), out.data_ptr(),
                              self._stream())
        return out

    def assemble(self, xi, yi, vals, n, normalize):
        if self.matrix is None or self.matrix.shape[0] != n:
            self.matrix = torch.empty((n, n), dtype=torch.float64, device=self.device)
        self.ctx.assemble_device(xi.numel(), xi.data_ptr(), yi.data_ptr(), vals.data_ptr(), n, normalize,
                                 self.matrix.data_ptr(), self._stream())
        return self.matrix


def broadcast_set(ctx, dset, src=0, group=None):
    """One compiled set on every rank: rank `src` passes its uploaded DeviceSet, the others None.  The records are
    compiled once (stemk_upload on `src`); the compiled set travels as ONE device-to-device broadcast (NCCL over
    NVLink) of stemk_set_export's byte string and becomes a set again through stemk_set_import -- instead of one
    host compile and one host->device copy per rank."""
    from .api import DeviceSet
    rank = dist.get_rank(group)
    dev = torch.device("cuda", ctx.device)
    nbytes = torch.tensor([dset.export_bytes() if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(nbytes, src, group=group)
    buf = torch.empty(int(nbytes.item()), dtype=torch.uint8, device=dev)
    if rank == src:
        dset.export_to(buf.data_ptr())          # returns when the bytes are in `buf`
    dist.broadcast(buf, src, group=group)
    torch.cuda.current_stream(dev).synchronize()
    return dset if rank == src else DeviceSet.import_from(ctx, buf.data_ptr(), buf.numel())


def record_keys(dset, stem=True, string=False):
    """Per-record size key of the work order (cost of the record against itself)."""
    v, e, length = dset.stats()
    k = np.zeros(len(v))
    if stem:
        k += 2.0 * v.astype(np.float64) * e
    if string:
        k += length.astype(np.float64) ** 2
    return k


# ----------------------------------------------------------------------------------------------- front end
class ShardedFold:
    """Base-pair probabilities of a batch of sequences over the ranks (the front end's McCaskill, fold.py / fold.cu).

    Sequences are independent: the batch is put in ONE global order, longest first (cost ~ L^3), dealt round-robin --
    rank r owns positions r, r+W, ... -- every rank folds its share on its own GPU and the pair lists travel to rank 0
    in one gather of objects (they are ragged).  No data-path collective besides that gather.  `fold_fn(seqs)` returns
    one (i, j, p) triple of arrays per sequence: fold.Folder(...).bpp(...).pairs on the GPUs, the oracle in the CPU
    test."""

    def __init__(self, lengths, rank, world):
        self.rank, self.world = rank, world
        self.order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
        self.mine = self.order[rank::world]

    def run(self, seqs, fold_fn):
        part = fold_fn([seqs[k] for k in self.mine])
        gathered = [None] * self.world if self.rank == 0 else None
        dist.gather_object([(int(k), tuple(np.asarray(a) for a in p)) for k, p in zip(self.mine, part)], gathered, dst=0)
        if self.rank != 0:
            return None
        out = [None] * len(seqs)
        for share in gathered:
            for k, p in share:
                out[k] = p
        return out
