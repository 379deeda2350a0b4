"""Per-source-line instruction / sample shares of one kernel launch in an .ncu-rep, grouped by function ranges
found from the source itself (so the grouping survives edits).  usage: ncu_regions.py rep kernel_regex [launch_skip] [topn]"""
import csv, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != '-': continue
    try: ln = int(r[0])
    except ValueError: continue
    ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); isamp = hdr.index("# Samples")
    g = lambda i: int(r[i]) if r[i] not in ('', '-') else 0
    data.append((cur, ln, r[1], g(ie), g(it), g(isamp)))
tot = sum(d[3] for d in data); tots = sum(d[5] for d in data); tt = sum(d[4] for d in data)
print("total warp inst %d thread inst %d (avg active %.1f) samples %d" % (tot, tt, tt / max(tot, 1), tots))
# function ranges from the .cu source on disk
import os
src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mirge3.0_b200", "csrc")
main = max(set(d[0] for d in data), key=lambda f: sum(d[3] for d in data if d[0] == f))
marks = []
try:
    for i, l in enumerate(open(os.path.join(src_dir, main)), 1):
        mm = re.match(r"^(?:template.*)?$", l)
        m2 = re.match(r"^(?:__device__|__global__|static|extern|trim_kernel|annotate_kernel)[^;]*?(\w+)\s*\(", l) or re.match(r"^(\w+)\(.*\{\s*$", l)
        if m2 and not l.startswith(" "): marks.append((i, m2.group(1)))
        m3 = re.match(r"^\s*// ?(?:---- )?(PHASE|REGION) (.*)$", l)
        if m3: marks.append((i, "  " + m3.group(2)[:30]))
except OSError:
    pass
marks.append((10 ** 9, "end"))
for (lo, name), (hi, _) in zip(marks, marks[1:]):
    ds = [d for d in data if d[0] == main and lo <= d[1] < hi]
    wi = sum(d[3] for d in ds); ti = sum(d[4] for d in ds); sm = sum(d[5] for d in ds)
    if wi: print("%-34s L%4d inst=%5.1f%% eff=%5.1f samp=%5.1f%%" % (name, lo, 100 * wi / tot, ti / max(wi, 1), 100 * sm / max(tots, 1)))
oth = [d for d in data if d[0] != main]
print("other files inst=%5.1f%%" % (100 * sum(d[3] for d in oth) / tot))
for d in sorted(data, key=lambda d: -d[3])[:topn]:
    print("%-10s %4d inst=%5.2f%% eff=%5.1f samp=%5.2f%% %s" % (d[0][:10], d[1], 100 * d[3] / tot, d[4] / max(d[3], 1), 100 * d[5] / max(tots, 1), d[2].strip()[:100]))
