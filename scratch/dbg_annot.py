import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import mirge_b200
from mirge_b200 import device as D, libraries as LB, manifoldAlign as MA, abi
from tests.test_gpu_annotate import make_libs, make_queries, oracle_annotate
from collections import Counter
dev = D.Device(0)
rng = np.random.default_rng(5)
libs = make_libs(rng)
seqs = make_queries(rng, libs, 6000)
ls = LB.LibrarySet.from_fasta_dict(dev, {k: (n, [s.encode() for s in v]) for k, (n, v) in libs.items()})
ks = MA.KeySet.from_strings(dev, seqs)
a_g, h_g = MA.annotate_keys(dev, ls, ks, True)
a_g = a_g.cpu().numpy(); h_g = h_g.cpu().numpy().view(np.uint64)
a_o, h_o = oracle_annotate(seqs, libs, True)
bad = np.nonzero((a_g != a_o) | (h_g != h_o))[0]
print("bad", bad.size, "of", len(seqs))
print("oracle rounds of bad:", Counter(a_o[bad].tolist()))
print("gpu rounds of bad:", Counter(a_g[bad].tolist()))
print("oracle rounds all:", Counter(a_o.tolist()))
print("len>32 among bad:", sum(len(seqs[i])>32 for i in bad), " among all:", sum(len(s)>32 for s in seqs))
mmc = Counter(int(h_o[i])>>56 for i in bad if a_o[i]!=255)
print("oracle mm of bad:", mmc)
for i in bad[:12]:
    s = seqs[i]; h = int(h_o[i]); r = (h>>28)&0xFFFFFFF; off = h&0xFFFFFFF
    if a_o[i] != 255:
        lib = libs[['mirna','hairpin','mature_trna','pre_trna','snorna','rrna','ncrna_others','mrna','mirna','spike-in'][a_o[i]]][1]
        ref = lib[r]
        q = s
        if a_o[i]==8: q = s[1:-2]
        if a_o[i]==3: q = s.rstrip('T')
        seg = ref[off:off+len(q)]
        mmpos = [j for j in range(len(q)) if q[j].upper()!=seg[j]]
        print(i, len(s), "o_round", a_o[i], "g_round", a_g[i], "ref", r, "off", off, "reflen", len(ref), "mmpos", mmpos, "gpu_hit %x" % int(h_g[i]))
    else:
        print(i, len(s), "oracle none; gpu round", a_g[i], "%x" % int(h_g[i]))
