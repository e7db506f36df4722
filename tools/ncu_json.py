"""Turn the ncu metrics pass over the stem-kernel launches of one Gram matrix into the JSON summary bench.py reads.

usage: python tools/ncu_json.py <metrics.csv> <n_pairs> <out.json> [source-note]

<metrics.csv> is the --csv --log-file output of
    ncu --metrics <METRICS below> --clock-control none -k regex:stem_ -s <skip the warm-up launches> -c <launches of one Gram> ...
over scripts/prof_stem.py N (one Gram matrix of N C3 records = N(N+1)/2 pairs per pass).  Sums are over the launches of
the matrix (one launch per size bucket); percentages are weighted by launch duration.
"""
import collections
import csv
import json
import sys

METRICS = ("gpu__time_duration.sum,smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,"
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,"
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,"
           "lts__t_sector_hit_rate.pct,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")


def main():
    if len(sys.argv) < 4:
        print(__doc__)
        print("METRICS=" + METRICS)
        return 2
    path, n_pairs, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        launches.setdefault(r[ix["ID"]], {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    L = list(launches.values())
    t = [l["gpu__time_duration.sum"] for l in L]   # ns
    T = sum(t)
    tot = lambda k: sum(l.get(k, 0.0) for l in L)
    avg = lambda k: sum(l.get(k, 0.0) * w for l, w in zip(L, t)) / T
    fma, add, mul = (tot("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k) for k in ("dfma", "dadd", "dmul"))
    res = {
        "source": sys.argv[4] if len(sys.argv) > 4 else path,
        "launches": len(L), "pairs": n_pairs, "kernel_ms": T * 1e-6, "pairs_per_s_under_ncu": n_pairs / (T * 1e-9),
        "dram_bytes_per_pair": (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) / n_pairs,
        "dram_read_bytes_per_pair": tot("dram__bytes_read.sum") / n_pairs,
        "dram_write_bytes_per_pair": tot("dram__bytes_write.sum") / n_pairs,
        "l2_bytes_per_pair": 32.0 * tot("lts__t_sectors.sum") / n_pairs,
        "executed_fp64_flop_per_pair": (2.0 * fma + add + mul) / n_pairs,
        "executed_fp64_instructions_per_pair": {"dfma": fma / n_pairs, "dadd": add / n_pairs, "dmul": mul / n_pairs},
        "warp_instructions_per_pair": tot("smsp__inst_executed.sum") / n_pairs,
        "lsu_wavefronts_per_pair": tot("l1tex__data_pipe_lsu_wavefronts.sum") / n_pairs,
        "shared_memory_wavefronts_per_pair": tot("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / n_pairs,
        "shared_memory_conflict_wavefronts_per_pair": tot("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / n_pairs,
        "issue_slots_busy": avg("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
        "lsu_data_pipe_of_peak": avg("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0,
        "shared_memory_wavefronts_of_peak": None,
        "fp64_pipe_active": avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
        "l2_hit_rate": avg("lts__t_sector_hit_rate.pct") / 100.0,
        "dram_throughput_of_peak": avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") / 100.0,
        "per_launch": [{"kernel": l["kernel"][:60], "ms": l["gpu__time_duration.sum"] * 1e-6,
                        "lsu_of_peak": l.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                        "l2_hit": l.get("lts__t_sector_hit_rate.pct")} for l in L],
    }
    if res["lsu_wavefronts_per_pair"]:
        res["shared_memory_wavefronts_of_peak"] = res["lsu_data_pipe_of_peak"] * res["shared_memory_wavefronts_per_pair"] / res["lsu_wavefronts_per_pair"]
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
