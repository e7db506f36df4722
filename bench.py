#!/usr/bin/env python
"""Benchmark of the Gram-matrix hot path (BASELINE.json metric: kernel-matrix pairs/s and GCUPS).

Workloads (`--workload`, the five BASELINE.json configs; SURVEY 8(d) concretises them):
  C3 (default, the one the driver runs)  stem_kernel_lite stem kernel (SuStemKernel, loop gap 0.2, beta 0.3, length
      band 10; threshold 0.01) over 2 000 synthetic ncRNA-like records of 150-300 nt with planted hairpins
      (stem_kernel_b200/synth.py, seed 20260003): 2 001 000 pairs per Gram matrix.  N > 1: weak scaling,
      records = round(2000*sqrt(N)) so that pairs per GPU stay constant.
  C4  the same kernel on a FIXED 10 000-record set (50 005 000 pairs): strong scaling over 2/4/8 GPUs.
  C5  rectangular 5 000 test x 10 000 train matrix (+ self terms and train diagonals), normalised, strong scaling;
      a sampled block feeds the vendored LIBSVM and must give the reference's predictions.
  C1  200 tRNA-like records, SuStemStrKernel, normalised: the whole matrix is compared with the reference and fed to
      LIBSVM cross-validation.
  C2  gap-weighted string kernel, 2 000 random sequences of 100 nt.
A "step" = one full normalised matrix (every pair, mirror / rectangular scatter, normalisation).

One process per GPU under torchrun; every rank holds the whole record set, owns a cost-balanced share of the pair
list, and rank 0 gathers the values over NCCL and assembles the matrix (stem_kernel_b200/sharded.py).

  value     whole-job pairs/s with the record set and pair lists already resident in HBM; timed per step with
            CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks
  e2e       the same metric through the host-buffer C ABI call a reference caller makes: stemk_upload (host
            compile + H2D of the flattened records) + stemk_gram / stemk_cross / stemk_pairs (H2D pair lists,
            kernels, D2H of the matrix)
  roofline  the stem kernel against the FP64 pipe: algorithmic flops (2*U_match + 3*U_bf + 3*U_skip
            per pair) / its own CUDA-event time, over the FP64 FMA peak measured live on the same GPU
            (MEASURED_PEAKS.json has no fp64 entry; its HBM figure is the denominator of roofline.hbm_view).
            traffic / executed flops / issue-slot figures are read from the committed ncu summary profiles/r02_*.json
  parity_check  rank 0's last timed matrix against the unmodified reference (oracle/_ref): >= 200 entries stratified
            over the (Lx, Ly) length buckets, both triangles, raw diagonals; the run exits non-zero above 1e-9
  cpu_baseline  the unmodified reference's kernel functors on the box's host cores, all threads, on a stratified
            sample of the same pairs

`--impl reference` times only that CPU arm, same metric and config.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TH = 0.01
BASE_N = 2000
METRIC = "kernel-matrix pairs/sec (stem kernel Gram matrix; GCUPS alongside)"
TOL = 1e-9
LEN_EDGES = (200, 250)     # length classes of the stratified samples: < 200, 200-249, >= 250 nt


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--records", type=int, default=0, help="override the record count (debugging only)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall budget of one CPU-baseline sample")
    ap.add_argument("--ref-variant", default="", choices=["", "o3"], help="reference arm: which build of oracle/_ref")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 / BPLA / naive-stem secondary figures")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- workloads
def workload(args, world):
    """name -> kernel kind, record counts, scaling and the config text."""
    from stem_kernel_b200 import _lib as L
    w = args.workload
    if w == "C3":
        n = args.records or int(round(BASE_N * math.sqrt(world)))
        return dict(name=w, kind=L.SU_STEM, cfg=3, n=n, n_test=0, scaling="weak", square=True, structure=True,
                    text="C3: stem_kernel_lite SuStemKernel Gram matrix, synthetic ncRNA-like records 150-300 nt with "
                         "planted hairpins, bp threshold 0.01, loop_gap 0.2, beta 0.3, len_band 10, normalised")
    if w == "C4":
        n = args.records or 10000
        return dict(name=w, kind=L.SU_STEM, cfg=4, n=n, n_test=0, scaling="strong", square=True, structure=True,
                    text="C4: SuStemKernel Gram matrix of a fixed set of ncRNA-like records 150-300 nt, cost-balanced "
                         "pair-list shards over the GPUs, normalised")
    if w == "C5":
        n = args.records or 10000
        return dict(name=w, kind=L.SU_STEM, cfg=5, n=n, n_test=n // 2, scaling="strong", square=False, structure=True,
                    text="C5: rectangular test x train SuStemKernel matrix (+ self terms and train diagonals), "
                         "normalised, feeding LIBSVM prediction")
    if w == "C1":
        n = args.records or 200
        return dict(name=w, kind=L.SU_STEM_STR, cfg=1, n=n, n_test=0, scaling="strong", square=True, structure=True,
                    text="C1: SuStemStrKernel Gram matrix of tRNA-like records (~75 nt), bp threshold 0.01, normalised")
    n = args.records or 2000
    return dict(name=w, kind=L.STR_SUBST, cfg=2, n=n, n_test=0, scaling="strong", square=True, structure=False,
                text="C2: lite StringKernel(gap 0.8, alpha 0.2) Gram matrix of random 100-nt sequences, normalised")


def make_records(W):
    from stem_kernel_b200 import synth
    train = synth.make_config(W["cfg"], W["n"])
    test = synth.make_config(W["cfg"], W["n_test"], offset=1000000) if W["n_test"] else None
    return train, test


def n_evals(W):
    n, nt = W["n"], W["n_test"]
    return n * (n + 1) // 2 if W["square"] else nt * n + nt + n


def workload_config(W, world, extra=None):
    c = {"workload": W["text"], "records": W["n"], "pairs": n_evals(W),
         "parallelism": f"pair-list sharding over {world} GPU(s), records replicated, one gather to rank 0"}
    if W["n_test"]:
        c["test_records"] = W["n_test"]
    if W["name"] == "C3":
        c["base_records_per_gpu_config"] = BASE_N
    if extra:
        c.update(extra)
    return c


def len_class(L):
    return np.digitize(np.asarray(L), LEN_EDGES)


# --------------------------------------------------------------------------------------------- reference arm
class RefPool:
    """The unmodified reference (oracle/_ref) on a pool of records: MData objects are built lazily by the
    reference's own constructor, pairs are evaluated by its kernel functors with n_th threads dealt round-robin
    like CalcTrainMatrix (common/kernel_matrix.cpp:42-56) and timed with a steady clock around the call."""

    def __init__(self, W, variant=""):
        if variant:
            os.environ["STEMK_REF_VARIANT"] = variant
        from oracle import refbind as R
        self.R = R
        self.W = W
        self.kernel = R.RefKernel(W["kind"])
        self.cache = {}

    def mdata(self, key, rec):
        if key not in self.cache:
            self.cache[key] = (self.R.RefMData.build(rec["rows"], rec["bp"], TH) if self.W["structure"]
                               else self.R.RefMData.seq_only(rec["rows"]))
        return self.cache[key]

    def pairs(self, recs_x, ix, recs_y, iy, threads, tag_x="a", tag_y="a"):
        """k(recs_x[ix[k]], recs_y[iy[k]]) for all k: (values, seconds)."""
        keys, ds = {}, []
        for tag, recs, idx in ((tag_x, recs_x, ix), (tag_y, recs_y, iy)):
            for i in idx:
                k = (tag, int(i))
                if k not in keys:
                    keys[k] = len(ds)
                    ds.append(self.mdata(k, recs[int(i)]))
        pi = np.array([keys[(tag_x, int(i))] for i in ix], dtype=np.int32)
        pj = np.array([keys[(tag_y, int(i))] for i in iy], dtype=np.int32)
        return self.kernel.pairs_timed(ds, pi, pj, n_th=threads)


def stratified_pairs(W, train, test, rng, pool_per_class, per_bucket):
    """Pair sample stratified over (Lx, Ly) length classes, drawn inside a pool of <= pool_per_class records per
    class (the reference's MData construction is front-end time and is not part of the metric, but it bounds how
    many distinct records a bounded sample can touch).  Returns a list of buckets
    dict(ix, iy, weight = pairs of the whole workload in this bucket) where x is the FIRST kernel argument."""
    Lx = np.array([len(r["rows"][0]) for r in train])
    cx = len_class(Lx)
    if W["square"]:
        Ly, cy, ysrc = Lx, cx, train
    else:
        Ly = np.array([len(r["rows"][0]) for r in test])
        cy, ysrc = len_class(Ly), test
    ncls = len(LEN_EDGES) + 1
    pool_x = [rng.permutation(np.nonzero(cx == a)[0])[:pool_per_class] for a in range(ncls)]
    pool_y = pool_x if W["square"] else [rng.permutation(np.nonzero(cy == a)[0])[:pool_per_class] for a in range(ncls)]
    buckets = []
    for a in range(ncls):
        for b in range(ncls):
            if W["square"] and b < a:
                continue
            na, nb = int(np.sum(cx == a)), int(np.sum(cy == b))
            weight = (na * (na + 1) // 2 if a == b else na * nb) if W["square"] else na * nb
            if weight == 0 or len(pool_x[a]) == 0 or len(pool_y[b]) == 0:
                continue
            m = min(per_bucket, weight)
            ix = pool_x[a][rng.integers(0, len(pool_x[a]), m)]
            iy = pool_y[b][rng.integers(0, len(pool_y[b]), m)]
            if W["square"]:     # the reference evaluates kernel(x_i, x_j) with i <= j (kernel_matrix.cpp:47-50)
                ix, iy = np.minimum(ix, iy), np.maximum(ix, iy)
            buckets.append(dict(a=a, b=b, ix=ix, iy=iy, weight=weight))
    return buckets, ysrc


def reference_rate(W, train, test, seconds, rng, threads, pool, pool_per_class=40):
    """pairs/s of the whole workload on the reference: every (Lx, Ly) bucket is timed on its own sample with all
    threads busy and the workload's time is extrapolated by the buckets' pair counts."""
    if W["name"] in ("C1", "C2"):
        # short records: one uniform bucket, sized for the budget
        n = W["n"]
        probe = 2000
        ix = rng.integers(0, n, probe); iy = rng.integers(0, n, probe)
        ix, iy = np.minimum(ix, iy), np.maximum(ix, iy)
        _, s0 = pool.pairs(train, ix, train, iy, threads)
        m = int(min(n * (n + 1) // 2, max(probe, probe * seconds / max(s0, 1e-6))))
        ix = rng.integers(0, n, m); iy = rng.integers(0, n, m)
        ix, iy = np.minimum(ix, iy), np.maximum(ix, iy)
        _, secs = pool.pairs(train, ix, train, iy, threads)
        return m / secs, m, secs, f"uniform sample of {m} pairs of the {n}-record set"
    buckets, ysrc = stratified_pairs(W, train, test, rng, pool_per_class, per_bucket=48)
    # calibration pass (also builds the pool's MData objects, outside the timed calls)
    rates = []
    for bk in buckets:
        _, s = pool.pairs(train, bk["ix"], ysrc, bk["iy"], threads, "a", "a" if W["square"] else "b")
        rates.append(len(bk["ix"]) / max(s, 1e-9))
    # size every bucket for an equal share of the wall budget, at least 200 and at most 4 000 pairs
    share = seconds / len(buckets)
    buckets2, _ = stratified_pairs(W, train, test, rng, pool_per_class, per_bucket=4000)
    total_pairs, total_secs, t_full = 0, 0.0, 0.0
    for bk, bk2, rate in zip(buckets, buckets2, rates):
        m = int(max(200, min(4000, rate * share)))
        ix, iy = bk2["ix"][:m], bk2["iy"][:m]
        _, s = pool.pairs(train, ix, ysrc, iy, threads, "a", "a" if W["square"] else "b")
        total_pairs += len(ix); total_secs += s
        t_full += bk["weight"] * s / len(ix)
    # (the self terms and train diagonals of the rectangular workload run at the buckets' mean rate)
    value = sum(b["weight"] for b in buckets) / t_full
    desc = (f"stratified: {len(buckets)} (Lx,Ly) length buckets (classes <200, 200-249, >=250 nt), {total_pairs} sampled "
            f"pairs in {total_secs:.1f} s drawn inside a pool of <= {pool_per_class} records per class, rate "
            f"extrapolated by the buckets' pair counts")
    return value, total_pairs, total_secs, desc


def run_reference(args, rank, world):
    if rank != 0:
        return
    W = workload(args, world)
    train, test = make_records(W)
    threads = os.cpu_count() or 1
    pool = RefPool(W, args.ref_variant)
    rng = np.random.default_rng(20260000 + W["cfg"])
    for _ in range(min(args.warmup, 1)):
        reference_rate(W, train, test, min(2.0, args.cpu_seconds), rng, threads, pool)
    vals, t_tot, n_tot, desc = [], 0.0, 0, ""
    for _ in range(args.steps):
        v, pairs, secs, desc = reference_rate(W, train, test, args.cpu_seconds, rng, threads, pool)
        vals.append(v); t_tot += secs; n_tot += pairs
    value = float(np.mean(vals))
    flags = "-O3 -march=x86-64-v3" if args.ref_variant == "o3" else "-g -O2 (autotools default)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1),
            "higher_is_better": True, "scaling": W["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(W, world, {"sampled_pairs_per_step": n_tot // max(args.steps, 1),
                                                 "note": "ms_per_step is the wall time of the sampled pairs; value "
                                                         "extrapolates them to the whole workload by bucket"}),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference",
                             "sample": desc + f"; unmodified reference sources compiled {flags}, kernel functors "
                                              "called with n_th = host threads like CalcTrainMatrix, wall clock"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- parity check
def parity_check(W, train, test, matrix, raw_diag_fn, rng, n_want=216):
    """Rank 0's last timed matrix against the unmodified reference on >= 200 entries stratified over the length
    buckets (+ both triangles of the square matrix, + raw diagonals).  matrix: numpy, normalised."""
    try:
        pool = RefPool(W)
    except Exception as ex:    # oracle/_ref missing on the box
        return {"unavailable": str(ex)}
    threads = os.cpu_count() or 1
    if W["name"] in ("C1", "C2"):
        n = W["n"]
        ix = rng.integers(0, n, n_want); iy = rng.integers(0, n, n_want)
        ix, iy = np.minimum(ix, iy), np.maximum(ix, iy)
        buckets, ysrc = [dict(ix=ix, iy=iy)], train
    else:
        buckets, ysrc = stratified_pairs(W, train, test, rng, pool_per_class=24,
                                         per_bucket=-(-n_want // (6 if W["square"] else 9)))
    ix = np.concatenate([b["ix"] for b in buckets]); iy = np.concatenate([b["iy"] for b in buckets])
    ty = "a" if W["square"] else "b"
    kxy, _ = pool.pairs(train, ix, ysrc, iy, threads, "a", ty)
    ux, uy = np.unique(ix), np.unique(iy)
    kxx, _ = pool.pairs(train, ux, train, ux, threads, "a", "a")
    kyy, _ = pool.pairs(ysrc, uy, ysrc, uy, threads, ty, ty)
    dx = dict(zip(ux.tolist(), kxx)); dy = dict(zip(uy.tolist(), kyy))
    with np.errstate(divide="ignore", invalid="ignore"):
        want = kxy / np.sqrt(np.array([dx[int(i)] for i in ix]) * np.array([dy[int(i)] for i in iy]))
    if W["square"]:
        want = np.where(ix == iy, 1.0, want)            # K_ii = 1 after normalisation (kernel_matrix.cpp:569)
        got = matrix[ix, iy]
        mirror_equal = bool(np.array_equal(got, matrix[iy, ix], equal_nan=True))
    else:
        got = matrix[iy, ix]                            # row = test record, column = train record
        mirror_equal = None
    nan_equal = bool(np.array_equal(np.isnan(got), np.isnan(want)))
    ok = ~np.isnan(want)
    err = float(np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok]))) if ok.any() else 0.0
    # raw (un-normalised) diagonals through the C ABI
    dsel = ux[:10]
    draw = raw_diag_fn(dsel)
    derr = float(np.max(np.abs(draw - np.array([dx[int(i)] for i in dsel])) / np.abs(np.array([dx[int(i)] for i in dsel]))))
    out = {"n": int(len(ix)), "max_rel_err": err, "nan_pattern_equal": nan_equal, "diagonals_checked": int(len(dsel)),
           "diag_max_rel_err": derr, "tolerance": TOL, "against": "oracle/_ref (unmodified reference, -g -O2)",
           "buckets": len(buckets)}
    if mirror_equal is not None:
        out["mirror_equal"] = mirror_equal
    out["ok"] = bool(err <= TOL and derr <= TOL and nan_equal and mirror_equal is not False)
    return out


def svm_block_check(W, train, test, ctx, rng, n_tr=160, n_te=80):
    """C5: a sampled train block and test block, both kernel matrices from the device and from the reference,
    through the vendored LIBSVM: identical predicted labels (BASELINE.md parity gate)."""
    from oracle import refbind as R
    from stem_kernel_b200 import hostlib
    tr = rng.choice(len(train), n_tr, replace=False); te = rng.choice(len(test), n_te, replace=False)
    y = np.array([train[int(i)]["label"] for i in tr], dtype=np.float64)
    mtr = hostlib.build_many([train[int(i)] for i in tr], TH); mte = hostlib.build_many([test[int(i)] for i in te], TH)
    dtr, dte = ctx.upload(mtr), ctx.upload(mte)
    Kg = ctx.gram(dtr, normalize=True)
    Cg, _ = ctx.cross(dte, dtr, normalize=True)
    k = R.RefKernel(W["kind"])
    rtr = [R.RefMData.build(train[int(i)]["rows"], train[int(i)]["bp"], TH) for i in tr]
    rte = [R.RefMData.build(test[int(i)]["rows"], test[int(i)]["bp"], TH) for i in te]
    Kr, _ = k.gram(rtr, normalize=True, n_th=os.cpu_count() or 1)
    Cr, _, _ = k.cross(rte, rtr, norm_test=True, normalize=True, n_th=os.cpu_count() or 1)
    pg, pr = R.svm_train_predict(Kg, y, Cg), R.svm_train_predict(Kr, y, Cr)
    dtr.free(); dte.free()
    return {"train_block": n_tr, "test_block": n_te, "labels_equal": bool(np.array_equal(pg, pr)),
            "max_rel_err_cross_block": float(np.max(np.abs(Cg - Cr) / np.abs(Cr)))}


def profile_summary(name):
    """Committed ncu summaries (profiles/r02_*.json, written by tools/ncu_json.py from the .ncu-rep files)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


# --------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from stem_kernel_b200 import _lib as L
    from stem_kernel_b200 import api, hostlib, sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL logs to stdout by default: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = workload(args, world)
    train, test = make_records(W)                      # every rank builds the same records (replicated input)
    n, nt = W["n"], W["n_test"]
    build = (lambda recs: hostlib.build_many(recs, TH)) if W["structure"] else \
            (lambda recs: [hostlib.MData.seq_only(r["rows"]) for r in recs])
    flat = hostlib.SeqSet(build(train))
    flat_t = hostlib.SeqSet(build(test)) if test else None
    params = L.make_params(W["kind"])
    ctx = api.Context(params, device=local_rank)
    dset = ctx.upload(flat)
    dtest = ctx.upload(flat_t) if flat_t else None
    stem, string = W["kind"] != L.STR_SUBST, W["kind"] in (L.SU_STEM_STR, L.STR_SUBST)
    keys = sharded.record_keys(dset, stem=stem, string=string)
    if W["square"]:
        be = sharded.GpuBackend(ctx, dset, dev)
        sg = sharded.ShardedGram(keys, rank, world, dev, be.compute, be.assemble)
        lists = [(dset, dset, sg.xi_host, sg.yi_host)]
        run_once = lambda: sg.run(normalize=True)
    else:
        be = sharded.GpuCrossBackend(ctx, dset, dtest, dev)
        sc = sharded.ShardedCross(sharded.record_keys(dtest, stem=stem, string=string), keys, rank, world, dev, be.compute)
        sets = {"train": dset, "test": dtest}
        lists = [(sets[wx], sets[wy], a, b) for wx, wy, a, b in sc.lists]
        run_once = lambda: sc.run(normalize=True)[0]
    n_pairs = sum(len(l[2]) for l in lists)

    # work model of this rank's share, reduced over ranks
    cells_r = flops_r = 0.0
    res_cells = 0.0
    for xs, ys, a, b in lists:
        mine = sharded.deal(len(a), rank, world)
        c_, f_ = ctx.pair_cost(xs, ys, a[mine], b[mine])
        cells_r += float(c_.sum()); flops_r += float(f_.sum())
        lx, ly = xs.stats()[2].astype(np.float64), ys.stats()[2].astype(np.float64)
        res_cells += float(np.sum(lx[a] * ly[b]))
    tot = torch.tensor([cells_r, flops_r], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    cells, flops = float(tot[0]), float(tot[1])
    my_pairs = sum(len(sharded.deal(len(l[2]), rank, world)) for l in lists)
    v_sz, e_sz, l_sz = dset.stats()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        with torch.cuda.stream(be.stream):
            return run_once()

    for _ in range(args.warmup):
        step()
    barrier()

    fp64_peak = ctx.fp64_peak(0.4)
    ctx.stats_reset()
    step_ms, last = [], None
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(be.stream):
                e0.record()
                last = run_once()
                e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            step_ms.append(float(t[0]))
    st = ctx.stats()
    total_ms = float(np.sum(step_ms))
    value = n_pairs * args.steps / (total_ms * 1e-3)
    kern_ms = (st["stem_ms"] + st["string_ms"]) / max(args.steps, 1)   # this rank's DP kernels, per step
    achieved = flops_r / (kern_ms * 1e-3) / 1e12
    launches_per_step = st["launches"] / max(args.steps, 1)
    last_host = last.cpu().numpy() if (rank == 0 and last is not None) else None

    # ---- end to end through the host-buffer C ABI (what a reference caller binds): upload + matrix, every step
    import ctypes as C
    desc = flat.desc()
    desc_t = flat_t.desc() if flat_t else None
    e2e_ms, e2e_upload_ms = [], []
    e2e_pinned = None
    mine_lists = [sharded.deal(len(l[2]), rank, world) for l in lists]
    for it in range(1 + args.steps):                       # first pass is a warm-up of the host path
        barrier()
        t0 = time.perf_counter()
        h, ht = C.c_void_p(), C.c_void_p()
        if world == 1:
            ctx._check(L.lib().stemk_upload(ctx.h, C.byref(desc), C.byref(h)))
            if desc_t is not None:
                ctx._check(L.lib().stemk_upload(ctx.h, C.byref(desc_t), C.byref(ht)))
        else:
            # the records are compiled ONCE, on rank 0; the compiled sets reach the other GPUs with one NCCL broadcast each
            e2e_sets = [sharded.broadcast_set(ctx, ctx.upload(flat) if rank == 0 else None)]
            h = e2e_sets[0].h
            if flat_t:
                e2e_sets.append(sharded.broadcast_set(ctx, ctx.upload(flat_t) if rank == 0 else None))
                ht = e2e_sets[1].h
        t_up = time.perf_counter() - t0
        if world == 1 and W["square"]:
            out = np.empty((n, n))
            ctx._check(L.lib().stemk_gram(ctx.h, h, 1, out.ctypes.data))
        elif world == 1:
            out = np.empty((nt, n)); selfv = np.empty(nt)
            ctx._check(L.lib().stemk_cross(ctx.h, ht, h, None, 0, 1, out.ctypes.data, selfv.ctypes.data))
        else:
            hs = {id(dset): h, id(dtest): ht}
            vals = []
            for (xs, ys, a, b), mine in zip(lists, mine_lists):
                xi_h = np.ascontiguousarray(a[mine]); yi_h = np.ascontiguousarray(b[mine])
                v = np.empty(len(mine))
                ctx._check(L.lib().stemk_pairs(ctx.h, hs[id(xs)], hs[id(ys)], len(mine), xi_h.ctypes.data,
                                               yi_h.ctypes.data, v.ctypes.data))
                vals.append(v)
            with torch.cuda.stream(be.stream):
                drv = sg if W["square"] else sc
                off = 0
                slabs = [drv.slab] if W["square"] else drv.slabs
                for v, sl_ in zip(vals, slabs):
                    drv.send[off:off + len(v)].copy_(torch.from_numpy(v))
                    off += sl_
                if rank == 0:
                    dist.gather(drv.send, list(drv.recv.unbind(0)), dst=0)
                    if W["square"]:
                        m = be.assemble(sg.xi_all, sg.yi_all, sharded.undeal(sg.recv, n_pairs), n, True)
                    else:
                        m = sc.finish(normalize=True)[0]
                    if e2e_pinned is None or e2e_pinned.shape != m.shape:
                        e2e_pinned = torch.empty(m.shape, dtype=torch.float64, pin_memory=True)
                    e2e_pinned.copy_(m, non_blocking=True)
                    torch.cuda.current_stream(dev).synchronize()
                    out = e2e_pinned.numpy()
                else:
                    dist.gather(drv.send, None, dst=0)
        if world == 1:
            L.lib().stemk_set_free(ctx.h, h)
            if ht:
                L.lib().stemk_set_free(ctx.h, ht)
        else:
            for s_ in e2e_sets:
                s_.free()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it > 0:
            e2e_ms.append(1e3 * float(dt[0]))
            e2e_upload_ms.append(1e3 * t_up)
    e2e_value = n_pairs * args.steps / (np.sum(e2e_ms) * 1e-3)
    set_bytes = dset.device_bytes() + (dtest.device_bytes() if dtest else 0)
    if world == 1:
        h2d = set_bytes + (8 * n_pairs if W["square"] else 0)        # records (+ the two uint32 index lists the caller's side builds)
        d2h = 8 * n * n if W["square"] else 8 * (nt * n + nt)
    else:
        h2d = (set_bytes if rank == 0 else 0) + 8 * my_pairs + 8 * my_pairs   # records (rank 0 only: the other ranks get them over NVLink), index lists, values back up for the gather
        d2h = 8 * my_pairs + ((8 * n * n if W["square"] else 8 * nt * n) if rank == 0 else 0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the timed run's own output against the unmodified reference
    rngp = np.random.default_rng(77 + world)

    def raw_diag(idx):
        return ctx.pairs(dset, dset, idx.astype(np.uint32), idx.astype(np.uint32))

    t0 = time.perf_counter()
    parity = parity_check(W, train, test, last_host, raw_diag, rngp)
    if W["name"] == "C5" and "unavailable" not in parity:
        parity["svm_predict"] = svm_block_check(W, train, test, ctx, rngp)
        parity["ok"] = bool(parity["ok"] and parity["svm_predict"]["labels_equal"])
    if W["name"] == "C1" and "unavailable" not in parity:
        from oracle import refbind as R
        pool = RefPool(W)
        Kr, _ = pool.kernel.gram([pool.mdata(("a", i), r) for i, r in enumerate(train)], normalize=True, n_th=os.cpu_count() or 1)
        yl = np.array([r["label"] for r in train], dtype=np.float64)
        parity["full_matrix_max_rel_err"] = float(np.max(np.abs(last_host - Kr) / np.abs(Kr)))
        parity["svm_cv_targets_equal"] = bool(np.array_equal(R.svm_cv(last_host, yl), R.svm_cv(Kr, yl)))
        parity["ok"] = bool(parity["ok"] and parity["full_matrix_max_rel_err"] <= TOL and parity["svm_cv_targets_equal"])
    parity["seconds"] = time.perf_counter() - t0

    # secondary workload, N=1 / C3 only: BASELINE.json configs[1] (C2: gap-weighted string kernel, 2 000 x 100 nt)
    secondary = None
    if world == 1 and W["name"] == "C3" and not args.records and not args.no_secondary:
        secondary = secondary_figures(local_rank)

    cpu = None
    if world == 1:
        try:
            pool = RefPool(W)
            v, pairs, secs, desc = reference_rate(W, train, test, args.cpu_seconds, np.random.default_rng(20260003),
                                                  os.cpu_count() or 1, pool)
            cpu = {"value": v, "unit": "pairs/s", "cores": os.cpu_count() or 1, "kind": "reference",
                   "sample": desc + "; unmodified reference sources compiled -g -O2 (autotools default), all host threads"}
            if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "o3", "libstemk_ref.so")):
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-variant", "o3",
                                    "--workload", W["name"], "--steps", "1", "--warmup", "0", "--cpu-seconds",
                                    str(min(6.0, args.cpu_seconds))] + (["--records", str(args.records)] if args.records else []),
                                   capture_output=True, text=True, timeout=600)
                try:
                    cpu["o3_value"] = json.loads(r.stdout.strip().splitlines()[-1])["value"]
                    cpu["o3_note"] = "courtesy row of BASELINE.md section 3: the same sources compiled -O3 -march=x86-64-v3 " \
                                     "(portable stand-in for -march=native: the library is built in another container)"
                except Exception:
                    pass
        except Exception as ex:  # oracle/_ref missing on the box
            cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    prof = profile_summary("r02_stem_ncu.json") if stem else None
    per_pair = (prof or {}).get("dram_bytes_per_pair")
    traffic = per_pair * my_pairs if per_pair else None
    roof = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak if fp64_peak else None,
            "traffic": traffic,
            "traffic_note": ("DRAM bytes of this rank's stem kernels per step = dram__bytes_read+write per pair of the ncu "
                             f"capture {prof.get('source')} x this rank's pairs" if prof else "no committed ncu summary found"),
            "kernel": "stem_fast_kernel" if stem else "string_pairs_kernel", "kernel_ms_per_step": kern_ms,
            "kernel_share_of_step": kern_ms / (total_ms / args.steps),
            "algorithmic_flop_per_pair": flops / max(n_pairs, 1),
            "peak_source": "FP64 FMA probe run live on this GPU (stemk_fp64_peak); MEASURED_PEAKS.json has no fp64 entry"}
    if prof:
        roof["executed"] = {k: prof.get(k) for k in ("executed_fp64_flop_per_pair", "warp_instructions_per_pair",
                                                     "issue_slots_busy", "shared_memory_wavefronts_of_peak",
                                                     "fp64_pipe_active", "l2_hit_rate", "dram_throughput_of_peak", "source")}
        if traffic and kern_ms:
            gbs = traffic / (kern_ms * 1e-3) / 1e9
            roof["hbm_view"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                "note": "traffic of the per-pair DP tables (not compulsory bytes) over the live kernel time; "
                                        "peak = measured hbm_gbs of MEASURED_PEAKS.json, else the 6650 GB/s fallback"}
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": W["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(W, world, {
            "l2": "256 MiB device write between timed steps; the per-step working set (record set + per-CTA DP "
                  "slabs) is several times the 126 MB L2",
            "dag_nodes_mean": float(v_sz.mean()), "dag_edges_mean": float(e_sz.mean())}),
        "gcups": cells * args.steps / (total_ms * 1e-3) / 1e9,
        "gcups_residue_cells": res_cells * args.steps / (total_ms * 1e-3) / 1e9,
        "alg_tflops": flops * args.steps / (total_ms * 1e-3) / 1e12,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": float(np.mean(e2e_ms)), "upload_ms_per_step": float(np.mean(e2e_upload_ms)),
                "path": ("stemk_upload + " + ("stemk_gram" if W["square"] else "stemk_cross") + " with host buffers") if world == 1 else
                        "stemk_upload on rank 0 + one NCCL broadcast of the compiled set, stemk_pairs with host buffers per rank, NCCL gather, device assemble, pinned D2H"},
        "gpu_launches": int(st["launches"]),
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof,
        "parity_check": parity,
        "cpu_baseline": cpu,
        "secondary": secondary,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if "ok" in parity and not parity["ok"]:
        raise SystemExit("parity_check failed: " + json.dumps(parity))


def secondary_figures(local_rank):
    """C2 (string kernel) and the built SURVEY 8(f) rows, one host-buffer call each."""
    from stem_kernel_b200 import _lib as L
    from stem_kernel_b200 import api, bpla, hostlib, nstem, synth
    recs2 = synth.make_config(2)
    md2 = [hostlib.MData.seq_only(r["rows"]) for r in recs2]
    ctx2 = api.Context(L.make_params(L.STR_SUBST), device=local_rank)
    ds2 = ctx2.upload(md2)
    n2 = len(md2)
    for _ in range(3):
        ctx2.gram(ds2)
    ctx2.stats_reset()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ctx2.gram(ds2, normalize=True)
    wall = (time.perf_counter() - t0) / reps
    st2 = ctx2.stats()
    p2 = n2 * (n2 + 1) // 2
    kms = st2["string_ms"] / reps
    secondary = {"workload": "C2: lite StringKernel(gap 0.8, alpha 0.2) Gram matrix, 2000 random sequences of 100 nt",
                 "pairs": p2, "kernel_ms": kms, "kernel_pairs_per_s": p2 / (kms * 1e-3),
                 "kernel_gcups": p2 * 1e4 / (kms * 1e-3) / 1e9,
                 "kernel_alg_tflops": 7.0 * p2 * 1e4 / (kms * 1e-3) / 1e12,
                 "e2e_pairs_per_s": p2 / wall, "e2e_ms": 1e3 * wall,
                 "e2e_note": "stemk_gram with host buffers (pair list built on the device, 32 MB matrix D2H)"}
    # the "next" rows of SURVEY 8(f) that are built (BPLA / local-alignment kernel, naive stem kernel): one
    # host-buffer call each on the C2 / C1 inputs, wall time (H2D and D2H inside)
    rng2 = np.random.default_rng(20260002)
    brecs = []
    for r in recs2:
        nb = len(r["rows"][0])
        a_, b_ = rng2.uniform(0, 0.6, nb), rng2.uniform(0, 0.4, nb)
        brecs.append(dict(rows=r["rows"], p_left=np.sqrt(a_), p_right=np.sqrt(b_), p_unpair=np.sqrt(np.maximum(0, 1 - a_ - b_))))
    bset = bpla.BplaSet(brecs)
    bxi, byi = np.triu_indices(n2)
    bpar = bpla.make_params()
    bpla.pairs(ctx2, bpar, bset, bset, bxi[:2000], byi[:2000])
    t0 = time.perf_counter()
    bpla.pairs(ctx2, bpar, bset, bset, bxi, byi)
    bw = time.perf_counter() - t0
    secondary["bpla"] = {"workload": "BPLAKernel (sum form, base-pairing profiles) on the C2 sequences, synthetic profiles",
                         "pairs": p2, "e2e_pairs_per_s": p2 / bw, "e2e_gcups": p2 * 1e4 / bw / 1e9}
    recs1 = synth.make_config(1)
    seqs1 = [r["rows"][0].lower() for r in recs1]
    nset = nstem.NstemSet(seqs1, [nstem.dense_bp(len(q), r["bp"][0], th=TH) for q, r in zip(seqs1, recs1)])
    nxi, nyi = np.triu_indices(len(seqs1))
    npar = nstem.make_params(bp_mode=1, bp_bound=TH)
    nstem.pairs(ctx2, npar, nset, nset, nxi[:148], nyi[:148])
    t0 = time.perf_counter()
    nstem.pairs(ctx2, npar, nset, nset, nxi, nyi)
    nw = time.perf_counter() - t0
    ln1 = np.array([len(q) for q in seqs1], dtype=np.float64)
    secondary["naive_stem"] = {"workload": "naive stem kernel (full_dp, probability tables, threshold 0.01) on the C1 records",
                               "pairs": len(nxi), "e2e_pairs_per_s": len(nxi) / nw,
                               "e2e_gcells": float(np.sum(ln1[nxi] ** 2 * ln1[nyi] ** 2) / 4.0) / nw / 1e9}
    # base-pair probabilities of the C3 sequences on the device (stemk_fold_bpp, host buffers in and
    # out); beside it the oracle restatement of the same recursions on one host core (the reference serialises its
    # ViennaRNA calls under a mutex, common/bpmatrix.cpp:152-155).  Parity with ViennaRNA itself is unpinned.
    from stem_kernel_b200 import fold
    from oracle import oraclebind as O
    seqs3 = [r["rows"][0] for r in synth.make_config(3)]
    fm = fold.default_model()
    folder = fold.Folder(ctx2)
    folder.bpp(seqs3, fm, cutoff=TH / 10)      # warm-up: the context keeps the DP scratch
    t0 = time.perf_counter()
    fr = folder.bpp(seqs3, fm, cutoff=TH / 10)
    fw = time.perf_counter() - t0
    t0 = time.perf_counter()
    worst = 0.0
    for k in range(0, len(seqs3), len(seqs3) // 8):
        want = O.fold_bpp(fm, seqs3[k])[0]
        i_, j_, p_ = fr.pairs[k]
        worst = max(worst, float(np.max(np.abs(p_ - want[i_, j_]) / want[i_, j_])) if len(p_) else 0.0)
    ow = (time.perf_counter() - t0) / 8
    nt3 = float(sum(len(q) for q in seqs3))
    secondary["fold"] = {"workload": "McCaskill base-pair probabilities of the C3 sequences (150-300 nt), stand-in loop model, cut-off 0.001",
                         "sequences": len(seqs3), "kernel_ms": fr.kernel_ms, "kernel_seqs_per_s": len(seqs3) / (fr.kernel_ms * 1e-3),
                         "e2e_seqs_per_s": len(seqs3) / fw, "e2e_nt_per_s": nt3 / fw,
                         "cpu_port_seqs_per_s_1core": 1.0 / ow, "parity_max_rel_err_8_sequences": worst,
                         "parity": "unpinned against ViennaRNA (absent); device = oracle restatement"}
    ctx2.close()
    return secondary


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
