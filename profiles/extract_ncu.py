#!/usr/bin/env python3
"""Selected metrics of ncu --set full captures as one CSV (what profiles/README.md quotes).

    python profiles/extract_ncu.py out.csv gpurun_out/r02c_rs_main.ncu-rep [more.ncu-rep ...]
"""
import csv
import os
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["capture"] + KEYS)
        for rep in reps:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(raw.splitlines()))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
                w.writerow([os.path.basename(rep).replace(".ncu-rep", "")] + [(d.get(k, "") + (" " + u.get(k, "") if u.get(k) and k != "Kernel Name" else "")).strip() for k in KEYS])


if __name__ == "__main__":
    main()
