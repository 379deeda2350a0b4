import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
keys=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio",
"sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__shared_mem_per_block_dynamic","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers",
"smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct",
"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","launch__grid_size","launch__block_size","launch__waves_per_multiprocessor"]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print("----")
    for k in keys:
        if k in d: print("%-90s %s %s"%(k, d[k], units[hdr.index(k)]))
