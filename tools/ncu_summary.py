"""Print the metrics that matter from an `ncu --page raw --csv` dump (one block per profiled launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
want += [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('-----')
    for i in idx:
        v = r[i]
        if 'stalled' in hdr[i]:
            try:
                if float(v.replace(',', '')) < 3: continue
            except ValueError: pass
        print(f"{hdr[i][:95]:95s} {v} {rows[1][i]}")
