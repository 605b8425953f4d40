"""Key metrics of an `ncu --page raw --csv` dump, one column per profiled launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w, units[i], [r[i][:40] for r in rows[2:]])
for i, hn in enumerate(hdr):
    if 'issue_stalled' in hn and 'per_warp_active' in hn:
        vals = [r[i] for r in rows[2:]]
        if any(float(v) > 2.0 for v in vals):
            print(hn.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), vals)
