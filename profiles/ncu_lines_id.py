"""Per-source-line view of one kernel of an ncu report: usage ncu_lines_id.py report.ncu-rep <kernel index> [top n] [inst|samp]
(sorted by stall samples, or by executed instructions with 'inst')."""
import csv, sys, subprocess
rep=sys.argv[1]; kid=sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 25
key = 3 if (len(sys.argv) > 4 and sys.argv[4] == "inst") else 5
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-id",":::"+str(int(kid)+1)],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; data=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<len(hdr) or r[2] != '-': continue
    try: ln=int(r[0])
    except: continue
    ie=hdr.index("Instructions Executed"); it=hdr.index("Thread Instructions Executed"); isamp=hdr.index("# Samples")
    g=lambda i: int(r[i]) if r[i] not in ('','-') else 0
    data.append((cur,ln,r[1],g(ie),g(it),g(isamp)))
tot=sum(d[3] for d in data); tots=sum(d[5] for d in data); tt=sum(d[4] for d in data)
print("total warp inst %d thread inst %d (avg active %.1f) samples %d"%(tot,tt,tt/max(tot,1),tots))
for d in sorted(data,key=lambda d:-d[key])[:topn]:
    print("%-10s %4d inst=%5.2f%% eff=%5.1f samp=%5.2f%% %s"%(d[0][:10], d[1], 100*d[3]/tot, d[4]/max(d[3],1), 100*d[5]/max(tots,1), d[2].strip()[:100]))
