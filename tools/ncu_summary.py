import csv,sys,subprocess
f=sys.argv[1]
out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','lts__t_bytes.sum','l1tex__t_bytes.sum','smsp__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_fp64.sum','smsp__inst_executed_pipe_fma.sum','smsp__inst_executed_pipe_alu.sum','smsp__inst_executed_pipe_lsu.sum','smsp__inst_executed_pipe_xu.sum']
for r in rows[2:]:
    print('kernel', r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr: print('   %-85s %s %s'%(w, r[hdr.index(w)], rows[1][hdr.index(w)]))
