"""Print the metrics that matter for an integer-multiply-bound kernel from an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_alu.', 'sm__inst_executed_pipe_lsu.', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput',
        'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate', 'l1tex__t_sector_pipe_lsu_mem_local_op_st_hit', 'lts__t_sector_hit_rate',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads',
        'sass__inst_executed_shared_stores', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct']
for vals in rows[2:]:
    print('#', vals[hdr.index('Kernel Name')][:80] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in want) or ('issue_stalled' in h and 'per_issue_active' in h):
            print(f"{h},{u},{v}")
