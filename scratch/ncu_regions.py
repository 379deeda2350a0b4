import csv, sys, subprocess, re
rep=sys.argv[1]; src=sys.argv[2]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; data={}
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<len(hdr) or r[2] != '-': continue
    try: ln=int(r[0])
    except: continue
    ie=hdr.index("Instructions Executed"); it=hdr.index("Thread Instructions Executed"); isamp=hdr.index("# Samples")
    g=lambda i: int(r[i]) if r[i] not in ('','-') else 0
    data[(cur,ln)]=(g(ie),g(it),g(isamp))
# regions = device functions / kernel in the source file, found by scanning for function starts
lines=open(src).read().split('\n')
base=src.split('/')[-1]
starts=[]
for i,l in enumerate(lines,1):
    m=re.match(r'^(?:template.*\n)?(?:__device__|__global__|static|extern).*?(\w+)\s*\(', l)
    if m and not l.startswith(' '): starts.append((i,m.group(1)))
    m2=re.match(r'^(\w+)\(const uint8_t \*__restrict__ fq', l)
    if m2: starts.append((i,m2.group(1)))
starts.sort()
def region(ln):
    name='(top)'
    for s,n in starts:
        if s<=ln: name=n
        else: break
    return name
agg={}
tot=[0,0,0]
for (f,ln),(a,b,c) in data.items():
    key=region(ln) if f==base else f
    x=agg.setdefault(key,[0,0,0]); x[0]+=a; x[1]+=b; x[2]+=c
    tot[0]+=a; tot[1]+=b; tot[2]+=c
print("region                         winst%   eff  samples%")
for k,(a,b,c) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print("%-30s %6.2f %5.1f %7.2f"%(k[:30],100*a/tot[0], b/max(a,1), 100*c/max(tot[2],1)))
