"""Summarise .ncu-rep captures (ncu --set full) as JSON: one object per kernel launch with the
metrics DESIGN.md cites.  Usage: python tools/ncu_summary.py out.json a.ncu-rep [b.ncu-rep ...]"""
import csv
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = dict(zip(head, row))
        u = dict(zip(head, units))
        o = {"Kernel Name": d.get("Kernel Name")}
        for k in KEYS:
            if k in d and d[k] != "":
                o[k] = f"{d[k]} {u.get(k, '')}".strip()
        res.append(o)
    return res


if __name__ == "__main__":
    dst, reps = sys.argv[1], sys.argv[2:]
    json.dump({os.path.splitext(os.path.basename(r))[0]: summarise(r) for r in reps}, open(dst, "w"), indent=1)
    print(open(dst).read()[:3000])
