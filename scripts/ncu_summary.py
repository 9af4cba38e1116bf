#!/usr/bin/env python3
"""Condense an ncu report (.ncu-rep, read here without a GPU) or a launch-list CSV into the small tables kept under profiles/.

  python scripts/ncu_summary.py full gpurun_out/prof_X.ncu-rep profiles/X_ncu_full.md
  python scripts/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.md
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full summary of `%s`\n\n" % rep)
        f.write("Captured with `ncu --set full --clock-control none --import-source on` (see scripts/gpu_profile.sh); "
                "times are cold-cache/serialised, never bench values.\n\n")
        f.write("| metric | unit | " + " | ".join(r[idx["Kernel Name"]].split("(")[0][-40:] + " #%s" % r[idx["ID"]] for r in data) + " |\n")
        f.write("|---|---|" + "---|"*len(data) + "\n")
        for k in KEYS:
            if k in idx:
                f.write("| %s | %s | " % (k, units[idx[k]]) + " | ".join(r[idx[k]] for r in data) + " |\n")
    print(open(out).read())


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    d = defaultdict(list)
    shape = {}
    for r in rows[1:]:
        d[r[ki]].append(float(r[vi].replace(",", "")))
        shape[r[ki]] = (r[gi], r[bi])
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write("# ncu launch list summary of `%s`\n\n" % path)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over bench steps; cold-cache serialised times: compare SHARES.\n\n")
        f.write("| kernel | launches | grid | block | total us | avg us | share |\n|---|---|---|---|---|---|---|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %s | %s | %.1f | %.1f | %.3f |\n" % (k.split("(")[0], len(v), shape[k][0], shape[k][1], sum(v)/1e3, sum(v)/len(v)/1e3, sum(v)/tot))
    print(open(out).read())


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
