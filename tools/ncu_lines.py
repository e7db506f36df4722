"""Summarise an .ncu-rep: headline metrics + warp-stall samples aggregated per CUDA source line.

usage: python tools/ncu_lines.py <file.ncu-rep> [top_n]   (needs -lineinfo builds and --import-source on)
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in KEYS:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(float(r[i] or 0), h[len(STALL):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
          if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
    st.sort(reverse=True)
    print("  stalls per issue: " + ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = {}
fname = ""
hdr2 = None
for row in csv.reader(io.StringIO(src)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr2 = row
        continue
    if hdr2 is None or row[0] in ("Function Name", "Kernel Name") or len(row) < len(hdr2):
        continue
    if row[2] != "-":      # SASS sub-rows; the CUDA line row carries the aggregate
        continue
    try:
        samples = int(row[hdr2.index("# Samples")])
        inst = int(row[hdr2.index("Instructions Executed")])
    except ValueError:
        continue
    stalls = {h: int(row[i] or 0) for i, h in enumerate(hdr2) if h.startswith("stall_") and "Not Issued" not in h}
    key = (fname, int(row[0]))
    if key in lines:
        lines[key][0] += samples
        lines[key][1] += inst
        for h, v in stalls.items():
            lines[key][3][h] = lines[key][3].get(h, 0) + v
    else:
        lines[key] = [samples, inst, row[1].strip(), stalls]
tot = sum(v[0] for v in lines.values()) or 1
print(f"== stall samples by source line (total {tot})")
for (f, ln), (s, inst, text, stalls) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    topst = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {100.0 * s / tot:5.1f}%  {f}:{ln:<4d} inst={inst:<9d} {' '.join(f'{k[6:]}={v}' for k, v in topst if v)}  | {text[:90]}")
