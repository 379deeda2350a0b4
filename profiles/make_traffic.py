"""DRAM traffic per read / per sequence of the hot kernels from one `ncu --set full` capture of a reduced bench run:

    python profiles/make_traffic.py gpurun_out/prof_X.ncu-rep READS SEQUENCES "capture description" > profiles/r2_traffic.json

Sums dram__bytes_read.sum + dram__bytes_write.sum over the launches of each kernel family in the report (the capture
holds one launch group of a READS-read pass) and divides by the reads (trim, collapse, tokeniser) or unique sequences
(annotate).  bench.py multiplies trim_kernel.bytes_per_read by the reads of one launch group for roofline.traffic."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

rep, reads, seqs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
desc = sys.argv[4] if len(sys.argv) > 4 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def fam(name):
    n = name.split("(")[0]
    if n.startswith("void "):
        n = n[5:]
    for key, f in (("trim_dp_kernel", "trim_kernel"), ("trim_kernel", "trim_kernel"), ("digest_tiles", "trim_kernel"),
                   ("collapse_list", "collapse"), ("assign_ids", "collapse"), ("tok_", "tokeniser"), ("annot", "annotate"),
                   ("drain", "drain")):
        if key in n:
            return f, n
    return None, n


agg = defaultdict(lambda: {"dram_read_bytes": 0, "dram_write_bytes": 0, "ms": 0.0, "per_launch": defaultdict(lambda: [0, 0, 0])})
for r in rows[2:]:
    f, n = fam(r[ki])
    if f is None:
        continue
    rd = float(r[ri].replace(",", "")) * scale.get(units[ri], 1)
    wr = float(r[wi].replace(",", "")) * scale.get(units[wi], 1)
    ms = float(r[ti].replace(",", "")) * tscale.get(units[ti], 1e-6)
    a = agg[f]
    a["dram_read_bytes"] += int(rd)
    a["dram_write_bytes"] += int(wr)
    a["ms"] += ms
    pl = a["per_launch"][n]
    pl[0] += int(rd)
    pl[1] += int(wr)
    pl[2] += 1
res = {"note": "DRAM traffic from one ncu --set full capture (dram__bytes_read.sum + dram__bytes_write.sum over the launches of "
               "a kernel family, cold cache, serialised); bench.py multiplies trim_kernel.bytes_per_read by the reads of one "
               "launch group to fill roofline.traffic", "capture": desc, "reads": reads, "sequences": seqs}
for f, a in agg.items():
    unit = seqs if f == "annotate" else reads
    d = {"dram_read_bytes": a["dram_read_bytes"], "dram_write_bytes": a["dram_write_bytes"], "ms_in_capture": round(a["ms"], 3),
         ("bytes_per_sequence" if f == "annotate" else "bytes_per_read"): round((a["dram_read_bytes"] + a["dram_write_bytes"]) / max(unit, 1), 1),
         "per_kernel": {k: {"read": v[0], "write": v[1], "launches": v[2]} for k, v in a["per_launch"].items()}}
    res[f] = d
print(json.dumps(res, indent=1))
