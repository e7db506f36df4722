#!/usr/bin/env python
"""Benchmark of the Gram-matrix hot path (BASELINE.json metric: kernel-matrix pairs/s and GCUPS).

Workload (config.workload = "C3"): BASELINE.json configs[2] -- stem_kernel_lite stem kernel (SuStemKernel, loop gap
0.2, beta 0.3, length band 10; threshold 0.01) over 2 000 synthetic ncRNA-like records of 150-300 nt with planted
hairpins (stem_kernel_b200/synth.py, seed 20260003): 2 001 000 pairs per Gram matrix.  It is the largest named
configuration that fits one GPU at a few seconds per step and the kernel the north star's target is quoted on.
A "step" = one full normalised Gram matrix (every pair i<=j, mirror, K_ij/sqrt(K_ii K_jj)).

N > 1 (weak scaling): one process per GPU under torchrun; records = round(2000*sqrt(N)) so that pairs per GPU stay
constant; every rank holds the whole record set, owns a cost-balanced share of the pair list, and rank 0 gathers
the values over NCCL and assembles the matrix (stem_kernel_b200/sharded.py).

  value     whole-job pairs/s with the record set and pair lists already resident in HBM; timed per step with
            CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks
  e2e       the same metric through the host-buffer C ABI call a reference caller makes: stemk_upload (host
            compile + H2D of the flattened records) + stemk_gram (H2D pair lists, kernels, D2H of the n x n matrix)
  roofline  the stem kernel against the FP64 pipe: algorithmic flops (2*U_match + 3*U_bf + 3*U_skip
            per pair) / its own CUDA-event time, over the FP64 FMA peak measured live on the same GPU
            (MEASURED_PEAKS.json has no fp64 entry; its HBM figure is used for the hbm sanity counter)
  cpu_baseline   the unmodified reference's threaded KernelMatrix::calculate (oracle/_ref) on the box's host cores,
            on a bounded random sub-matrix of the same records

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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=0, help="override the record count (debugging only)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall budget of one CPU-baseline sample")
    return ap.parse_args()


def n_records(world, override):
    return override if override else int(round(BASE_N * math.sqrt(world)))


def workload_config(n, world, extra=None):
    c = {"workload": "C3: stem_kernel_lite SuStemKernel Gram matrix, synthetic ncRNA-like records 150-300 nt "
                     "with planted hairpins, bp threshold 0.01, loop_gap 0.2, beta 0.3, len_band 10, normalised",
         "records": n, "pairs": n * (n + 1) // 2, "base_records_per_gpu_config": BASE_N,
         "parallelism": f"pair-list sharding over {world} GPU(s), records replicated, one gather to rank 0"}
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------- reference arm
def reference_sample(recs, seconds, rng, threads):
    """Times the UNMODIFIED reference (oracle/_ref: KernelMatrix::calculate with n_th threads,
    common/kernel_matrix.cpp:485-575) on random sub-matrices of `recs`; returns pairs/s and a description."""
    from oracle import refbind as R
    n = len(recs)
    m = min(n, max(2 * threads // 3, 16))
    k = R.RefKernel(R.SU_STEM)
    built = {}

    def ref_of(i):
        if i not in built:
            built[i] = R.RefMData.build(recs[i]["rows"], recs[i]["bp"], TH)
        return built[i]

    # calibrate on a small sub-matrix, then size the sample for the wall budget
    idx = rng.choice(n, size=m, replace=False)
    _, secs = k.gram([ref_of(int(i)) for i in idx], normalize=True, n_th=threads)
    rate = (m * (m + 1) / 2) / max(secs, 1e-6)
    m2 = int(min(n, max(m, math.sqrt(2.0 * rate * seconds))))
    idx = rng.choice(n, size=m2, replace=False)
    _, secs = k.gram([ref_of(int(i)) for i in idx], normalize=True, n_th=threads)
    pairs = m2 * (m2 + 1) // 2
    return pairs / secs, pairs, m2, secs


def cells_of(recs_sizes, xi, yi):
    v = recs_sizes.astype(np.float64)
    return float(np.sum(v[xi] * v[yi]))


def run_reference(args, rank, world):
    if rank != 0:
        return
    from stem_kernel_b200 import synth
    n = n_records(world, args.records)
    recs = synth.make_config(3, n)
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(20260003)
    for _ in range(min(args.warmup, 1)):
        reference_sample(recs, min(2.0, args.cpu_seconds), rng, threads)
    vals, t_tot, desc = [], 0.0, ""
    for _ in range(args.steps):
        v, pairs, m, secs = reference_sample(recs, args.cpu_seconds, rng, threads)
        vals.append(v); t_tot += secs
        desc = f"random {m}-record sub-matrix ({pairs} pairs) of the {n}-record set per step"
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, world),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference",
                             "sample": desc + "; unmodified reference sources compiled -g -O2 (autotools default), "
                                              "KernelMatrix::calculate with n_th = host threads, wall clock"},
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


# --------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from stem_kernel_b200 import _lib as L
    from stem_kernel_b200 import api, hostlib, sharded, synth

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

    n = n_records(world, args.records)
    recs = synth.make_config(3, n)                    # every rank builds the same records (replicated input)
    md = hostlib.build_many(recs, TH)
    flat = hostlib.SeqSet(md)
    params = L.make_params(L.SU_STEM)
    ctx = api.Context(params, device=local_rank)
    dset = ctx.upload(flat)
    be = sharded.GpuBackend(ctx, dset, dev)
    keys = sharded.record_keys(dset)
    sg = sharded.ShardedGram(keys, rank, world, dev, be.compute, be.assemble)
    n_pairs = sg.n_pairs

    # work model of this rank's share, reduced over ranks
    mine = sharded.deal(n_pairs, rank, world)
    cells_r, flops_r = ctx.pair_cost(dset, dset, sg.xi_host[mine], sg.yi_host[mine])
    tot = torch.tensor([cells_r.sum(), flops_r.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    cells, flops = float(tot[0]), float(tot[1])
    my_flops = float(flops_r.sum())
    v_sz, e_sz, l_sz = dset.stats()
    res_cells = float(np.sum(l_sz.astype(np.float64)[sg.xi_host] * l_sz.astype(np.float64)[sg.yi_host]))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        with torch.cuda.stream(be.stream):
            return sg.run(normalize=True)

    for _ in range(args.warmup):
        step()
    barrier()

    fp64_peak = ctx.fp64_peak(0.4)
    ctx.stats_reset()
    step_ms = []
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(be.stream):
                e0.record()
                sg.run(normalize=True)
                e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            step_ms.append(float(t[0]))
    st = ctx.stats()
    total_ms = float(np.sum(step_ms))
    value = n_pairs * args.steps / (total_ms * 1e-3)
    kern_ms = st["stem_ms"] / max(args.steps, 1)          # this rank's stem kernel, average launch duration
    achieved = my_flops / (kern_ms * 1e-3) / 1e12
    launches_per_step = st["launches"] / max(args.steps, 1)

    # ---- end to end through the host-buffer C ABI (what a reference caller binds): upload + gram, every step
    desc = flat.desc()
    import ctypes as C
    e2e_ms, e2e_upload_ms = [], []
    n_loc = len(mine)
    for it in range(1 + args.steps):                       # first pass is a warm-up of the host path
        barrier()
        t0 = time.perf_counter()
        h = C.c_void_p()
        ctx._check(L.lib().stemk_upload(ctx.h, C.byref(desc), C.byref(h)))
        t_up = time.perf_counter() - t0
        if world == 1:
            out = np.empty((n, n))
            ctx._check(L.lib().stemk_gram(ctx.h, h, 1, out.ctypes.data))
        else:
            xi_h = np.ascontiguousarray(sg.xi_host[mine]); yi_h = np.ascontiguousarray(sg.yi_host[mine])
            vals = np.empty(n_loc)
            ctx._check(L.lib().stemk_pairs(ctx.h, h, h, n_loc, xi_h.ctypes.data, yi_h.ctypes.data, vals.ctypes.data))
            with torch.cuda.stream(be.stream):
                send = torch.zeros(sg.slab, dtype=torch.float64, device=dev)
                send[:n_loc].copy_(torch.from_numpy(vals))
                if rank == 0:
                    dist.gather(send, list(sg.recv.unbind(0)), dst=0)
                    m = be.assemble(sg.xi_all, sg.yi_all, sharded.undeal(sg.recv, n_pairs), n, True)
                    out = m.cpu().numpy()
                else:
                    dist.gather(send, None, dst=0)
        L.lib().stemk_set_free(ctx.h, h)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it > 0:
            e2e_ms.append(1e3 * float(dt[0]))
            e2e_upload_ms.append(1e3 * t_up)
    e2e_value = n_pairs * args.steps / (np.sum(e2e_ms) * 1e-3)
    set_bytes = dset.device_bytes()
    h2d = set_bytes + 8 * (n_pairs if world == 1 else n_loc)           # records + two uint32 index lists
    d2h = 8 * n * n if (world == 1 or rank == 0) else 0
    if world > 1:
        h2d += 8 * n_loc
        d2h += 8 * n_loc

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # secondary workload, N=1 only: BASELINE.json configs[1] (C2: gap-weighted string kernel, 2 000 x 100 nt)
    secondary = None
    if world == 1 and not args.records:
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
                     "e2e_pairs_per_s": p2 / wall, "e2e_note": "stemk_gram with host buffers (pair lists H2D, 32 MB matrix D2H)"}
        # the "next" rows of SURVEY 8(f) that are built (BPLA / local-alignment kernel, naive stem kernel): one
        # host-buffer call each on the C2 / C1 inputs, wall time (H2D and D2H inside)
        from stem_kernel_b200 import bpla, nstem
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
        ctx2.close()

    cpu = None
    if world == 1:
        try:
            v, pairs, m, secs = reference_sample(recs, args.cpu_seconds, np.random.default_rng(20260003),
                                                 os.cpu_count() or 1)
            cpu = {"value": v, "unit": "pairs/s", "cores": os.cpu_count() or 1, "kind": "reference",
                   "sample": f"random {m}-record sub-matrix ({pairs} pairs, {secs:.1f} s wall) of the same {n} records; "
                             "unmodified reference KernelMatrix::calculate, all host threads"}
        except Exception as ex:  # oracle/_ref missing on the box
            cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, world, {
            "l2": "256 MiB device write between timed steps; the per-step working set (record set + per-CTA DP "
                  "slabs) is several times the 126 MB L2",
            "dag_nodes_mean": float(v_sz.mean()), "dag_edges_mean": float(e_sz.mean()),
            "deal_imbalance": sharded.imbalance(sharded.stem_cost_proxy(v_sz, e_sz, l_sz, sg.xi_host, sg.yi_host), world)}),
        "gcups": cells * args.steps / (total_ms * 1e-3) / 1e9,
        "gcups_residue_cells": res_cells * args.steps / (total_ms * 1e-3) / 1e9,
        "alg_tflops": flops * args.steps / (total_ms * 1e-3) / 1e12,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": float(np.mean(e2e_ms)), "upload_ms_per_step": float(np.mean(e2e_upload_ms)),
                "path": "stemk_upload + stemk_gram with host buffers" if world == 1 else
                        "stemk_upload + stemk_pairs with host buffers per rank, NCCL gather, stemk_assemble_device, D2H"},
        "gpu_launches": int(st["launches"]),
        "gpu_launches_per_step": launches_per_step,
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak if fp64_peak else None,
                     # DRAM bytes of the stem kernels of one step: 4.44 MB per pair (read 3.27 + write 1.17) from the ncu
                     # capture profiles/r01_stem_final_traffic_n256.csv, scaled by this rank's pair count
                     "traffic": 4.44e6 * len(mine),
                     "traffic_note": "per step, summed over the step's bucket launches like `achieved`; the per-pair DP "
                                     "table (G0 slab, ~1.2 MB) lives in L2/HBM by design, see DESIGN.md 5.1",
                     "kernel": "stem_fast_kernel", "kernel_ms_per_launch": kern_ms,
                     # what the kernel actually saturates (ncu capture, not live): the recursion is a gather-accumulate
                     # through shared memory, so the instruction issue slots and the shared-memory pipe fill up long
                     # before the FP64 pipe does
                     "practical_bounds": {"issue_slots_busy": 0.647, "shared_memory_wavefronts_of_peak": 0.501,
                                          "fp64_pipe_active": 0.168, "dram_throughput_of_peak": 0.182,
                                          "source": "profiles/r01_stem_final_ncu_summary.txt (ncu --set full, 256 C3 records)"},
                     "kernel_share_of_step": kern_ms / (total_ms / args.steps),
                     "peak_source": "FP64 FMA probe run live on this GPU (stemk_fp64_peak); MEASURED_PEAKS.json has no "
                                    "fp64 entry",
                     # the same kernel against the HBM roofline: its DRAM traffic (ncu capture above) over its live
                     # launch time, against the measured copy bandwidth of MEASURED_PEAKS.json
                     "hbm_view": {"bound": "hbm", "achieved": 4.44e6 * len(mine) / (kern_ms * 1e-3) / 1e9 if kern_ms else None,
                                  "peak": hbm_peak, "unit": "GB/s",
                                  "frac": (4.44e6 * len(mine) / (kern_ms * 1e-3) / 1e9 / hbm_peak) if kern_ms and hbm_peak else None,
                                  "note": "traffic of the per-pair DP tables (not compulsory bytes); peak = measured "
                                          "hbm_gbs of MEASURED_PEAKS.json, else the 6650 GB/s fallback"},
                     "hbm_sanity": {"set_bytes": int(set_bytes), "peak_gbs": hbm_peak,
                                    "note": "compulsory HBM traffic is << 1 B/flop on this path (SURVEY 8(d))"}},
        "cpu_baseline": cpu,
        "secondary": secondary,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
