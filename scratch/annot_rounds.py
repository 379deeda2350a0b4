"""Per-round timing of the annotation kernel (unfused launches) on the C2 workload, split by key class."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mirge_b200
from mirge_b200 import device as D, libraries as LB, manifoldAlign as MA, synth
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = D.Device(0)
libs = synth.make_libraries(mrna_count=100_000)
lset = LB.LibrarySet.from_fasta_dict(dev, libs.fasta_dict())
eng = D.DigestEngine(dev, synth.trim_config_for(2, "head"))
fq = synth.ReadGenerator(libs, synth.CONFIGS[2], dev.tdev).fastq(n_reads)
table = D.CollapseTable(dev, min_keys=1 << 24)
eng.digest_device(fq, table, 2048 << 20)
keys = MA.KeySet.from_table(table)
print("keys", keys.n)
for fused in (True, False):
    for rep in range(2):
        dev.timing = True
        dev.timer_totals()
        annot, hit = MA.annotate_keys(dev, lset, keys, False, fused=fused, ordered=(fused and rep == 1))
        t = dev.timer_totals()
    print("fused" if fused else "per round", {k: round(v[1], 3) for k, v in sorted(t.items())}, "total", round(sum(v[1] for v in t.values()), 3))
a = annot.cpu()
import numpy as np
print("annotated per round", np.bincount(a.numpy()[a.numpy() != 0xFF], minlength=9).tolist())
# key length classes
lens = (table.arena[table.key_ref[:keys.n].long()] & 0xFFFF).cpu().numpy()
print("len<26:", int((lens < 26).sum()), "26..40:", int(((lens >= 26) & (lens <= 40)).sum()), ">40:", int((lens > 40).sum()))
