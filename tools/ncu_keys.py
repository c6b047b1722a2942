"""Print the handful of ncu raw-page metrics used when reading a capture: python tools/ncu_keys.py file.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed.sum', 'launch__occupancy_limit_shared_mem', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum']
for val in rows[2:]:
    for i, h in enumerate(hdr):
        if h in keys or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h) or \
           ('sm__inst_executed_pipe_' in h and h.endswith('.avg.pct_of_peak_sustained_active')):
            if val[i] not in ('0', '0.0', ''):
                print(f"{h:75s} {units[i]:12s} {val[i]}")
