#!/usr/bin/env python
"""Print the headline metrics of one `ncu --set full` report (first kernel in it)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, u, v = r[0], r[1], r[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed', 'sm__throughput.avg.pct',
        'dram__throughput.avg.pct', 'issue_stalled']
for i, n in enumerate(h):
    if any(w in n for w in want):
        if 'issue_stalled' in n and 'ratio' not in n:
            continue
        if v[i] in ('', '0'):
            continue
        print(n, u[i], v[i])
