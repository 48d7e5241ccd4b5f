"""Print selected metrics of every kernel in an `ncu --page raw --csv` dump:  python tools/ncu_pick.py raw.csv [metric ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = sys.argv[2:] or ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
                        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
                        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
                        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
                        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
                        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
                        "dram__bytes_read.sum", "dram__bytes_write.sum"]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", r[ik][:70])
    for w in want:
        if w in d:
            print(f"   {w:90s} {d[w]}")
