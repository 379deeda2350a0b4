#!/usr/bin/env python
"""Generates the ``ref_*`` fixtures of this directory by running UNMODIFIED reference code.

    python tests/golden/make_reference_golden.py            (needs /root/reference; writes tests/golden/ref_case1/)

What runs from the reference (imported from /root/reference, not copied):
  * ``mirge.libs.digest.baking``           -- stipulate(), the chunk fan-out, the worker ``cutadapt(n)`` with its
    per-modifier (HEAD) counting, the qiagen split trick and UMIParser, the parent merge, both UMI levels,
    the sample x sequence matrix, the three counter dicts, ``_umiCounts.csv`` and ``.trim.collapse.fa``
    (digest.py:59-375);
  * ``mirge.libs.manifoldAlign.bwtAlign``  -- the round driver: length / annotFlag masks, the ``T{3,}$`` query
    rewrite of round 3, FASTA writing, command assembly, SAM parsing with "later lines overwrite earlier
    ones" (manifoldAlign.py:12-146);
  * the annotFlag split + ``to_csv`` of ``mirge/__main__.py:164-173`` (restated here in the two lines it is);
  * ``mirge.libs.summary.summarize``       -- annotation.report.csv, miR.Counts.csv, miR.RPM.csv
    (summary.py:677-1290).

What cannot run here and is stood in for (no network, nothing vendored): the third-party tools.
  * ``cutadapt``, ``dnaio``, ``xopen`` are stand-in modules (tests/golden/standins.py) exposing the API subset
    digest.py touches, implemented with the oracle's restatement of their semantics; ``Bio`` is an empty
    stub (nothing of it is reached).
  * ``bowtie`` and ``bowtie-inspect`` are small scripts written into a temporary directory that
    ``args.bowtie_path`` points at.  The stand-in bowtie answers with the oracle's definition of the
    valid hit set (oracle/pyoracle.py: end-to-end, ungapped, forward strand, -n / -v / -e 70 policy, best
    stratum) and prints the canonical pick LAST, so the reference's "last SAM line wins" parser lands on
    it.  bowtie's own choice among equally good alignments is therefore NOT pinned by these fixtures
    (DESIGN.md section 2); everything the reference itself does around bowtie is.
So the fixtures pin everything the reference does itself and leave the third-party arithmetic (cutadapt's
alignment, bowtie's search) to the oracle's restatement, which has no real install to be checked against here.

pandas here is 3.x, the reference was written against 1.x; the script fails loudly if the reference code
does not run under it.
"""
import argparse
import gzip
import os
import shutil
import stat
import sys
import tempfile
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REFERENCE = Path(os.environ.get("MIRGE_REFERENCE", "/root/reference"))

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

CASE = HERE / "ref_case1"
ORG = "synth"
DB = "miRBase"
SAMPLES = ["sampleA", "sampleB"]


FAKE_BOWTIE = r'''#!%(python)s
"""Stand-in for bowtie 1.x used ONLY to drive the reference's round loop (see make_reference_golden.py)."""
import sys
sys.path.insert(0, %(root)r)
from oracle import pyoracle as po

argv = sys.argv[1:]
opts = {"-n": None, "-v": None, "-5": 0, "-3": 0, "--threads": 1}
flags = set()
pos = []
i = 0
while i < len(argv):
    a = argv[i]
    if a in opts:
        opts[a] = int(argv[i + 1]); i += 2
    elif a.startswith("-"):
        flags.add(a); i += 1
    else:
        pos.append(a); i += 1
index, fasta = pos[0], pos[1]
lib = po.read_fasta(open(index + ".fa").read())
if opts["-v"] is not None:
    pol = po.RoundPolicy(0, opts["-v"], opts["-v"], trim5=opts["-5"], trim3=opts["-3"])
else:
    pol = po.RoundPolicy(28, opts["-n"] if opts["-n"] is not None else 2, 2, trim5=opts["-5"], trim3=opts["-3"])
out = sys.stdout
out.write("@HD\tVN:1.0\tSO:unsorted\n")
for n, s in zip(lib.names, lib.seqs):
    out.write("@SQ\tSN:%%s\tLN:%%d\n" %% (n, len(s)))
out.write("@PG\tID:Bowtie\tVN:stand-in\tCL:\"%%s\"\n" %% " ".join(sys.argv))
name = None
for line in open(fasta):
    line = line.rstrip("\n")
    if line.startswith(">"):
        name = line[1:]
        continue
    q = line[pol.trim5: max(len(line) - pol.trim3, pol.trim5)]
    hs = po.hits(q, lib, pol)
    if not hs:
        out.write("%%s\t4\t*\t0\t0\t*\t*\t0\t0\t%%s\t%%s\tXM:i:0\n" %% (name, q, "I" * len(q)))
        continue
    best = min(h[0] for h in hs)
    hs = sorted((h[0], h[1], h[2]) for h in hs if h[0] == best)
    if "-a" not in flags:
        hs = hs[:1]
    for mm, r, off in reversed(hs):  # canonical pick (minimum) last: the reference keeps the last line
        out.write(po.sam_line(name, q, lib.names[r], lib.seqs[r], off, pol) + "\n")
'''

FAKE_INSPECT = r'''#!%(python)s
"""Stand-in for ``bowtie-inspect -n <index>``: reference names of <index>.fa, one per line."""
import sys
index = [a for a in sys.argv[1:] if not a.startswith("-")][-1]
for line in open(index + ".fa"):
    if line.startswith(">"):
        print(line[1:].rstrip("\n"))
'''


def make_libraries(rng):
    B = np.array(list("ACGT"))

    def rnd(n, lo, hi):
        return ["".join(rng.choice(B, rng.integers(lo, hi + 1))) for _ in range(n)]

    mir = rnd(48, 20, 24)  # summarize() draws a 5 x 8 tile map of the top 40 miRNAs: needs >= 40 names after merging
    libs = {
        "mirna": (["hsa-miR-%d-%dp" % (i // 2 + 1, 5 if i % 2 == 0 else 3) for i in range(48)], mir),
        "hairpin": (["hsa-mir-%d" % (i + 1) for i in range(24)], rnd(24, 70, 100)),
        "mature_trna": (["tRNA-%d" % i for i in range(12)], [s + "CCA" for s in rnd(12, 70, 80)]),
        "pre_trna": (["pre-tRNA-%d" % i for i in range(12)], rnd(12, 90, 110)),
        "snorna": (["SNORD%d" % i for i in range(15)], rnd(15, 60, 120)),
        "rrna": (["RNA5S%d" % i for i in range(3)], rnd(3, 120, 400)),
        "ncrna_others": (["ncRNA%d" % i for i in range(20)], rnd(20, 100, 300)),
        "mrna": (["NM_%06d" % i for i in range(20)], rnd(20, 300, 600)),
        "spike-in": (["spike%d" % i for i in range(5)], rnd(5, 22, 22)),
    }
    # hairpins carry their two mature arms
    hp = libs["hairpin"][1]
    for i in range(24):
        a, b = mir[2 * i], mir[2 * i + 1]
        hp[i] = hp[i][:5] + a + hp[i][5 + len(a): 45] + b + hp[i][45 + len(b):]
    return libs


def make_fastq(rng, libs, n, sample_idx):
    """Reads = library fragments (exact, isomiR-like, with substitutions, polyT tails) + adapter, two length
    classes, with a few N / low-quality tails; HEAD counting makes every pre-adapter read a key as well."""
    from tests.util import ILL

    B = np.array(list("ACGT"))
    keys = list(libs)
    w = np.array([8, 1, 1.5, 1, 1, 1, 1, 1, 0.5])
    w = w / w.sum()
    out = []
    for i in range(n):
        k = keys[int(rng.choice(len(keys), p=w))]
        names, seqs = libs[k]
        j = int(rng.zipf(1.6) + sample_idx) % len(seqs)
        ref = seqs[j]
        if k == "mirna":
            ins = ref
            r = rng.random()
            if r < 0.25:  # isomiR: shifted ends (templated from nothing: random flanks)
                ins = str(rng.choice(B)) + ref + "".join(rng.choice(B, 2))
            elif r < 0.35:
                ins = ref[1:-1]
            elif r < 0.40:
                p = int(rng.integers(len(ref)))
                ins = ref[:p] + str(rng.choice(B)) + ref[p + 1:]
        else:
            L = int(rng.integers(18, 40))
            a = int(rng.integers(0, max(1, len(ref) - L + 1)))
            ins = ref[a:a + L]
            if k == "pre_trna" and rng.random() < 0.7:
                ins = ins[:25] + "TTTT"
            if rng.random() < 0.2:
                p = int(rng.integers(len(ins)))
                ins = ins[:p] + str(rng.choice(B)) + ins[p + 1:]
        if rng.random() < 0.05:
            ins = "".join(rng.choice(B, int(rng.integers(16, 35))))
        s = (ins + ILL + "".join(rng.choice(B, 50)))[:50]
        if rng.random() < 0.02:
            p = int(rng.integers(len(s)))
            s = s[:p] + "N" + s[p + 1:]
        q = np.clip(rng.integers(28, 41, len(s)) - (np.arange(len(s)) > 40) * rng.integers(0, 30), 2, 41)
        out.append("@%s.%d\n%s\n+\n%s\n" % (SAMPLES[sample_idx], i, s, "".join(chr(33 + int(v)) for v in q)))
    return "".join(out).encode()


ILLUMINA = "TGGAATTCTCGGGTGCCAAGGAACTCCAG"
QIA_INNER = "AACTGTAGGCACCATCAAT"


def reference_args(libdir=None, bindir=None, **kw):
    """The args namespace fields the reference code reads (mirge/libs/parse.py defaults unless overridden)."""
    a = argparse.Namespace(threads=1, bowtie_path=str(bindir) if bindir else None, bowtieVersion="True", quiet=True,
                           organism_name=ORG, libraries_path=str(libdir) if libdir else None, spikeIn=True, bam_out=False,
                           tRNA_frag=False, crThreshold="0.1", gff_out=False, AtoI=False, isoform_entropy=False,
                           novel_miRNA=False,
                           adapters=[("back", ILLUMINA)], error_rate=0.12, overlap=3, indels=True, match_adapter_wildcards=True,
                           match_read_wildcards=False, times=1, action="trim", nextseq_trim=None, quality_cutoff="10",
                           phred64=33, trim_n=False, cut=[], minimum_length=16, uniq_mol_ids=None, qiagenumi=False,
                           umiDedup=False, tcf_out=False, fasta=False, buffer_size=60000, cutadaptVersion=("3", "1"))
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def digest_cases(rng):
    """Extra digest-only cases: (directory name, args overrides, FASTQ bytes per sample)."""
    B = np.array(list("ACGT"))

    def reads(n, build, L=60, lowq_tail=True, tag="r"):
        out = []
        for i in range(n):
            s = build(i)[:L]
            q = np.clip(rng.integers(25, 41, len(s)) - (np.arange(len(s)) > L - 12) * rng.integers(0, 32) * lowq_tail, 2, 41)
            out.append("@%s.%d\n%s\n+\n%s\n" % (tag, i, s, "".join(chr(33 + int(v)) for v in q)))
        return "".join(out).encode()

    pool = ["".join(rng.choice(B, rng.integers(17, 28))) for _ in range(60)]

    def insert():
        r = rng.random()
        if r < 0.85:
            return pool[int(rng.zipf(1.4)) % len(pool)]
        if r < 0.93:
            return "".join(rng.choice(B, rng.integers(5, 15)))  # shorter than -m after trimming
        return "".join(rng.choice(B, rng.integers(16, 40)))

    def noisy(ad, p=0.03):
        return "".join(str(rng.choice(B)) if rng.random() < p else c for c in ad)

    def rnd(k):
        return "".join(rng.choice(B, k))

    umi_pool = [rnd(8) for _ in range(12)]  # few UMIs: PCR duplicates collapse at the first level

    def ill4n(i):  # 4N + insert + 4N + adapter (docs/source/quick_start.md "-umi 4,4")
        u = umi_pool[int(rng.integers(len(umi_pool)))]
        return u[:4] + insert() + u[4:] + noisy(ILLUMINA) + rnd(40)

    def qia(i):  # insert + inner adapter + 12-nt UMI + outer adapter (--qiagenumi -umi 0,12)
        ins = insert()
        if rng.random() < 0.05:
            ins = ins + ins[:6] + ins  # the trimmed read occurs twice: exercises str.split(trimmed)[1]
        tail = noisy(QIA_INNER) + rnd(12) + "AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC"
        if rng.random() < 0.1:
            tail = tail[: int(rng.integers(3, 25))]
        return ins + tail + rnd(30)

    def nxt(i):  # poly-G tails, N ends, cuts
        s = insert() + noisy(ILLUMINA) + "G" * int(rng.integers(0, 30)) + rnd(30)
        if rng.random() < 0.2:
            s = "N" * int(rng.integers(1, 3)) + s
        return s

    FRONT = "GTTCAGAGTTCTACAGTCCGACGATC"  # the reference's "illumina" alias for -g (mirge/__main__.py:74-76)

    def both(i):  # 5' adapter remnant + insert + 3' adapter: -g and -a together, two removal rounds, no indels
        s = insert() + noisy(ILLUMINA, 0.05) + rnd(30)
        r = rng.random()
        if r < 0.35:
            s = noisy(FRONT[int(rng.integers(0, 20)):], 0.04) + s
        elif r < 0.4:
            s = rnd(int(rng.integers(1, 4))) + FRONT + s  # 5' adapter not at the very start
        return s

    return [
        ("ref_case2_umi", dict(uniq_mol_ids="4,4", umiDedup=False, tcf_out=True), [reads(1500, ill4n, tag="u%d" % k) for k in range(2)]),
        ("ref_case3_umi_dedup", dict(uniq_mol_ids="4,4", umiDedup=True), [reads(1500, ill4n, tag="d")]),
        ("ref_case4_qiagen", dict(adapters=[("back", QIA_INNER)], uniq_mol_ids="0,12", qiagenumi=True, umiDedup=True, tcf_out=True),
         [reads(1500, qia, L=75, tag="q%d" % k) for k in range(2)]),
        ("ref_case5_nextseq_cuts", dict(nextseq_trim=20, quality_cutoff="5,15", trim_n=True, cut=[2, -3], times=2, minimum_length=14),
         [reads(1500, nxt, L=75, tag="n")]),
        ("ref_case6_front_back_noindels", dict(adapters=[("back", ILLUMINA), ("front", FRONT)], indels=False, times=2, minimum_length=18,
                                               tcf_out=True),
         [reads(1200, both, L=75, tag="f%d" % k) for k in range(3)]),
    ]


def run_digest_case(baking, name, overrides, fastqs):
    """Reference baking() on the case's FASTQ files; writes inputs, the DataFrame and the counters."""
    import json

    case = HERE / name
    if case.exists():
        shutil.rmtree(case)
    case.mkdir()
    samples = ["s%d" % (i + 1) for i in range(len(fastqs))]
    files = []
    for sname, data in zip(samples, fastqs):
        f = case / (sname + ".fastq")
        f.write_bytes(data)
        files.append(str(f))
    args = reference_args(**overrides)
    with tempfile.TemporaryDirectory() as tmp:
        df, src, trc, tru = baking(args, files, samples, Path(tmp))            # digest.py:105
        df.sort_index(kind="stable").to_csv(case / "complete_set.csv")         # single-sample order is arbitrary: sorted
        for extra in sorted(os.listdir(tmp)):
            if extra.endswith("_umiCounts.csv") or extra.endswith(".trim.collapse.fa"):
                shutil.copy(Path(tmp) / extra, case / extra)
    json.dump({"args": {k: v for k, v in overrides.items()}, "samples": samples, "sampleReadCounts": src, "trimmedReadCounts": trc,
               "trimmedReadCountsUnique": tru, "dtypes": {c: str(t) for c, t in df.dtypes.items()}},
              open(case / "counters.json", "w"), indent=1, sort_keys=True)
    print("wrote", case, "rows", len(df), src, trc, tru)


def main():
    if not REFERENCE.exists():
        sys.exit("reference checkout %s not found" % REFERENCE)
    from tests.golden import standins

    standins.install(REFERENCE)
    from mirge.libs.digest import baking
    from mirge.libs.manifoldAlign import bwtAlign  # the reference's own code
    from mirge.libs.summary import summarize

    rng = np.random.default_rng(20260117)
    libs = make_libraries(rng)
    if CASE.exists():
        shutil.rmtree(CASE)
    libdir = CASE / "lib"
    idx = libdir / ORG / "index.Libs"
    idx.mkdir(parents=True)
    (libdir / ORG / "annotation.Libs").mkdir()
    suffix = {"mirna": "_mirna_" + DB, "hairpin": "_hairpin_" + DB, "mature_trna": "_mature_trna", "pre_trna": "_pre_trna",
              "snorna": "_snorna", "rrna": "_rrna", "ncrna_others": "_ncrna_others", "mrna": "_mrna", "spike-in": "_spike-in"}
    for k, (names, seqs) in libs.items():
        with open(idx / (ORG + suffix[k] + ".fa"), "w") as f:
            for n, s in zip(names, seqs):
                f.write(">%s\n%s\n" % (n, s))
    # merges file: miR-1-5p and miR-2-5p are reported under one name (summary.py:705-714)
    (libdir / ORG / "annotation.Libs" / ("%s_merges_%s.csv" % (ORG, DB))).write_text(
        "hsa-miR-1-5p/2-5p,hsa-miR-1-5p,hsa-miR-2-5p\nhsa-miR-7-3p/9-3p,hsa-miR-7-3p,hsa-miR-9-3p\n")
    fastqs = [make_fastq(rng, libs, 2500, si) for si in range(2)]
    for name, data in zip(SAMPLES, fastqs):
        with gzip.GzipFile(CASE / (name + ".fastq.gz"), "wb", mtime=0) as f:
            f.write(data)

    files = [str(CASE / (name + ".fastq.gz")) for name in SAMPLES]

    with tempfile.TemporaryDirectory() as tmp:
        bindir = Path(tmp) / "bin"
        bindir.mkdir()
        for fname, text in (("bowtie", FAKE_BOWTIE), ("bowtie-inspect", FAKE_INSPECT)):
            p = bindir / fname
            p.write_text(text % {"python": sys.executable, "root": str(ROOT)})
            p.chmod(p.stat().st_mode | stat.S_IEXEC)
        work = Path(tmp) / "work"
        work.mkdir()
        args = reference_args(libdir, bindir, quality_cutoff="20")
        df, src, trc, tru = baking(args, files, SAMPLES, work)               # digest.py:105
        out = bwtAlign(args, df, work, DB)                                   # manifoldAlign.py:68
        pdMapped = out[out.annotFlag.eq(1)]                                  # __main__.py:164
        pdUnmapped = out[out.annotFlag.eq(0)]                                # __main__.py:165
        summarize(args, work, DB, SAMPLES, pdMapped, src, trc, tru)          # __main__.py:166
        pdMapped.to_csv(work / "mapped.csv")                                 # __main__.py:172
        pdUnmapped.to_csv(work / "unmapped.csv")                             # __main__.py:173
        for f in ("mapped.csv", "unmapped.csv", "annotation.report.csv", "miR.Counts.csv", "miR.RPM.csv"):
            shutil.copy(work / f, CASE / f)
        # the per-round SAM files of -bam / -trf (manifoldAlign.py:20-62): bwtAlign alone, on a fresh table
        work2 = Path(tmp) / "work_sam"
        work2.mkdir()
        args2 = reference_args(libdir, bindir, quality_cutoff="20", bam_out=True, tRNA_frag=True)
        df2, *_ = baking(args2, files, SAMPLES, work2)
        out2 = bwtAlign(args2, df2, work2, DB)
        assert out2.equals(out)
        (CASE / "sam").mkdir()
        for f in sorted(os.listdir(work2)):
            if f.endswith(".sam"):
                shutil.copy(work2 / f, CASE / "sam" / f)
    n_map, n_un = len(pdMapped), len(pdUnmapped)
    (CASE / "README.txt").write_text(
        "Generated by tests/golden/make_reference_golden.py (pandas %s) from the unmodified reference's bwtAlign,\n"
        "annotFlag split and summarize; third-party tools stood in for as the script header explains.\n"
        "Trim settings: -a illumina -q 20, count_mode=head, -spk, crThreshold 0.1.  %d mapped / %d unmapped sequences.\n"
        % (pd.__version__, n_map, n_un))
    print("wrote", CASE, "mapped", n_map, "unmapped", n_un)
    print((CASE / "annotation.report.csv").read_text())
    for name, overrides, fq in digest_cases(rng):
        run_digest_case(baking, name, overrides, fq)


if __name__ == "__main__":
    main()
