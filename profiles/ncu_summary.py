#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of counters the design decisions rest on.
usage: python profiles/ncu_summary.py report.ncu-rep [--md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ_%"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_pipe_%"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lsu_wavefronts_shared"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "lsu_wavefronts_global"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_sel"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        print("### %s  (id %s)" % (name[:110], d.get("ID", "?")))
        for k, short in KEYS:
            if k in d:
                print("  %-24s %s %s" % (short, d[k], u[k]))
        print()


if __name__ == "__main__":
    main()
