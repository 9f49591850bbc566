"""Summarise an ncu report (--page raw --csv) into the few numbers the design notes quote.
    python tools/ncu_summary.py gpurun_out/full_x.ncu-rep [...]"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]; units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('#', rep)
    for r in rows[2:]:
        print('--', r[idx['Kernel Name']][:60], 'grid', r[idx['Grid Size']], 'block', r[idx['Block Size']])
        for w in WANT:
            if w in idx:
                print('    %-82s %s %s' % (w, r[idx[w]], units[idx[w]]))
