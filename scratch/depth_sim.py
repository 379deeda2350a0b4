"""How many columns back does cutadapt's traceback path of a cost>=2 candidate go before its cost drops to <=1?
(decides how many Myers columns the trim kernel must keep at hand to resolve candidates without the slow path)"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mirge_b200
from mirge_b200 import synth
from oracle import pyoracle as po
AD = synth.ILLUMINA; m = len(AD)
acc = [int(0.12 * i) if i >= 3 else -1 for i in range(m + 1)]
libs = synth.make_libraries(scale=0.05, mrna_count=50)
fq = bytes(synth.ReadGenerator(libs, synth.CONFIGS[2], "cpu").fastq(6000).numpy())
lines = fq.split(b"\n")
hist = collections.Counter(); ncand = 0; reads_with = 0; nreads = 0
worst = collections.Counter()
for r in range(0, len(lines) - 3, 4):
    seq = lines[r + 1].decode(); qual = lines[r + 3].decode()
    stop = po.nextseq_trim_index(seq, qual, 20, 33)
    seq, qual = seq[:stop], qual[:stop]
    s, e = po.quality_trim_index(qual, 0, 20, 33)
    read = seq[s:e]; n = len(read); nreads += 1
    if n == 0: continue
    # DP with back pointers: step kind 0 diag-match,1 mismatch,2 ins,3 del
    cost = np.zeros((m + 1, n + 1), dtype=np.int32); kind = np.zeros((m + 1, n + 1), dtype=np.int8)
    cost[:, 0] = np.arange(m + 1)
    for j in range(1, n + 1):
        for i in range(1, m + 1):
            if AD[i - 1] == read[j - 1]:
                cost[i, j] = cost[i - 1, j - 1]; kind[i, j] = 0
            else:
                cd, cdel, cins = cost[i - 1, j - 1] + 1, cost[i, j - 1] + 1, cost[i - 1, j] + 1
                if cd <= cdel and cd <= cins: cost[i, j] = cd; kind[i, j] = 1
                elif cins <= cdel: cost[i, j] = cins; kind[i, j] = 2
                else: cost[i, j] = cdel; kind[i, j] = 3
    exact = [j for j in range(1, n + 1) if cost[m, j] == 0]
    jmax = exact[0] if exact else n
    cands = [(m, j) for j in range(1, jmax) if cost[m, j] <= acc[m]]
    if not exact: cands += [(i, n) for i in range(1, m + 1) if cost[i, n] <= acc[i]]
    w = 0
    for (i, j) in cands:
        if cost[i, j] < 2: continue
        ncand += 1
        a, b = i, j
        while a > 0 and b > 0 and cost[a, b] > 1:
            k = kind[a, b]
            if k in (0, 1): a -= 1; b -= 1
            elif k == 2: a -= 1
            else: b -= 1
        d = j - b
        hist[d] += 1; w = max(w, d)
    if any(cost[i, j] >= 2 for (i, j) in cands): reads_with += 1; worst[w] += 1
print("reads", nreads, "with cost>=2 candidates", reads_with, "candidates", ncand)
print("depth histogram (columns back until cost<=1):", sorted(hist.items()))
print("per-read worst depth:", sorted(worst.items()))
