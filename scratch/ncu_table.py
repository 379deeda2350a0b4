import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
cols=[("gpu__time_duration.sum","us"),("smsp__inst_executed.sum","winst"),("smsp__thread_inst_executed_per_inst_executed.ratio","eff"),
("sm__warps_active.avg.pct_of_peak_sustained_active","warps%"),("smsp__issue_active.avg.pct_of_peak_sustained_active","issue%"),
("lts__t_sector_hit_rate.pct","L2hit"),("l1tex__t_sector_hit_rate.pct","L1hit"),("dram__bytes_read.sum","dramRd"),
("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","longsb"),("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","wait"),
("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","shortsb"),("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","branch"),
("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","noinst"),("launch__registers_per_thread","regs")]
print(" ".join("%9s"%c[1] for c in cols))
for r in rows[2:]:
    d=dict(zip(hdr,r))
    vals=[]
    for k,_ in cols:
        v=d.get(k,"")
        try: v="%.4g"%float(v.replace(",",""))
        except: pass
        vals.append(v)
    print(" ".join("%9s"%v for v in vals))
