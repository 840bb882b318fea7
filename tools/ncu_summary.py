#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, no GPU): python tools/ncu_summary.py rep1 rep2 ..."""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__bytes_read.sum.per_second',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__issue_active.max.pct_of_peak_sustained_active',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active',
 'launch__registers_per_thread','launch__grid_size','launch__block_size','sm__cycles_elapsed.avg.per_second',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio','smsp__average_warps_issue_stalled_misc_per_issue_active.ratio']
cols = {}
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            cols.setdefault(k, {})[rep] = (rows[2][i], units[i])
names = sys.argv[1:]
print("%-95s" % "metric" + "".join("%16s" % n.split('/')[-1].replace('.ncu-rep','')[-15:] for n in names))
for k in KEYS:
    if k in cols:
        u = next(iter(cols[k].values()))[1]
        print("%-95s" % (k[-80:] + " [" + u + "]") + "".join("%16s" % cols[k].get(n, ("-",))[0][:15] for n in names))
