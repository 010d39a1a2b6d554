"""One-line summary of the first kernel in an .ncu-rep (raw page).  usage: ncu_summary.py <rep> [more reps]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print(rep.split('/')[-1], v[h.index('Kernel Name')][:60])
        for k in KEYS:
            if k in h: print("   %-82s %s %s" % (k, v[h.index(k)], u[h.index(k)]))
