"""The oracle against files written by the UNMODIFIED reference (tests/golden/ref_case1, made by
tests/golden/make_reference_golden.py from /root/reference): bwtAlign's round driver, the annotFlag split +
to_csv, and summarize()'s annotation.report.csv / miR.Counts.csv.  CPU only; the GPU twin of this test is
tests/test_gpu_report.py."""
import gzip
import os

import pytest

from mirge_b200 import params as P
from oracle import pyoracle as po
from tests.util import py_params

CASE = os.path.join(os.path.dirname(__file__), "golden", "ref_case1")
SAMPLES = ["sampleA", "sampleB"]
ORG, DB = "synth", "miRBase"
SUFFIX = {"mirna": "_mirna_" + DB, "hairpin": "_hairpin_" + DB, "mature_trna": "_mature_trna", "pre_trna": "_pre_trna",
          "snorna": "_snorna", "rrna": "_rrna", "ncrna_others": "_ncrna_others", "mrna": "_mrna", "spike-in": "_spike-in"}


def golden(name):
    with open(os.path.join(CASE, name)) as f:
        return f.read()


def case_config():
    return P.TrimConfig(adapters=[("back", "TGGAATTCTCGGGTGCCAAGGAACTCCAG")], quality_cutoff="20", count_mode="head")


def load_case_libraries():
    libs = {}
    for k, suf in SUFFIX.items():
        with open(os.path.join(CASE, "lib", ORG, "index.Libs", ORG + suf + ".fa")) as f:
            libs[k] = po.read_fasta(f.read())
    return libs


@pytest.fixture(scope="module")
def oracle_run():
    p = py_params(case_config())
    counts, src, trc, tru = {}, {}, {}, {}
    for j, s in enumerate(SAMPLES):
        with gzip.open(os.path.join(CASE, s + ".fastq.gz"), "rb") as f:
            d = po.digest_sample(f.read(), p)
        src[s], trc[s], tru[s] = d.count, d.trimmed, len(d.table)
        for k, c in d.table.items():
            counts.setdefault(k, [0, 0])[j] = c
    libs = load_case_libraries()
    annot = po.annotate(sorted(counts), libs, spike_in=True)
    return counts, annot, libs, src, trc, tru


def test_mapped_and_unmapped_csv_match_reference(oracle_run):
    counts, annot, libs, *_ = oracle_run
    mapped, unmapped = po.split_tables(annot, counts, SAMPLES, spike_in=True)
    assert mapped == golden("mapped.csv")
    assert unmapped == golden("unmapped.csv")


def test_annotation_report_matches_reference(oracle_run):
    counts, annot, libs, src, trc, tru = oracle_run
    merges = golden(os.path.join("lib", ORG, "annotation.Libs", "%s_merges_%s.csv" % (ORG, DB)))
    report, mir = po.summarize_counts(annot, counts, SAMPLES, src, trc, tru, merges, libs["mirna"].names, 0.1, True)
    assert po.report_csv(report, SAMPLES, True) == golden("annotation.report.csv")
    assert po.mir_counts_csv(mir, SAMPLES) == golden("miR.Counts.csv")
    assert po.mir_rpm_csv(mir, SAMPLES) == golden("miR.RPM.csv")


# ---- digest-only cases: the reference's baking() (worker, parent merge, UMI levels, matrix, counters, side files)

DIGEST_CASES = ["ref_case2_umi", "ref_case3_umi_dedup", "ref_case4_qiagen", "ref_case5_nextseq_cuts", "ref_case6_front_back_noindels"]


def load_digest_case(name):
    import json

    from tests.util import make_args

    d = os.path.join(os.path.dirname(__file__), "golden", name)
    meta = json.load(open(os.path.join(d, "counters.json")))
    ov = dict(meta["args"])
    if "adapters" in ov:
        ov["adapters"] = [tuple(a) for a in ov["adapters"]]
    args = make_args(**ov)
    files = [os.path.join(d, s + ".fastq") for s in meta["samples"]]
    return d, meta, args, files


def sorted_lines(path, skip_header=True):
    with open(path) as f:
        lines = f.read().splitlines()
    return sorted(lines[1:] if skip_header else lines)


def tcf_pairs(path):
    """(count, sequence) pairs of a .trim.collapse.fa and the check that headers are numbered by descending count."""
    with open(path) as f:
        lines = f.read().splitlines()
    pairs = []
    for i in range(0, len(lines), 2):
        n, c = lines[i][4:].split("_")
        assert int(n) == i // 2 + 1
        pairs.append((int(c), lines[i + 1]))
    assert [c for c, _ in pairs] == sorted((c for c, _ in pairs), reverse=True)
    return sorted(pairs)


@pytest.mark.parametrize("name", DIGEST_CASES)
def test_oracle_digest_matches_reference_baking(name):
    d, meta, args, files = load_digest_case(name)
    p = py_params(P.TrimConfig.from_args(args, "head"))
    tables, src, trc, tru = [], {}, {}, {}
    for s, f in zip(meta["samples"], files):
        dg = po.digest_sample(open(f, "rb").read(), p, umi_dedup=bool(args.umiDedup), buffer_size=60000)
        tables.append(dg.table)
        src[s], trc[s], tru[s] = dg.count, dg.trimmed, len(dg.table)
        umi_csv = os.path.join(d, s + "_umiCounts.csv")
        if os.path.exists(umi_csv):
            assert sorted("%s,%s,%d" % r for r in dg.umi_rows) == sorted_lines(umi_csv)
        tcf = os.path.join(d, s + ".trim.collapse.fa")
        if os.path.exists(tcf):
            assert sorted((c, k) for k, c in dg.table.items()) == tcf_pairs(tcf)
    assert src == meta["sampleReadCounts"] and trc == meta["trimmedReadCounts"] and tru == meta["trimmedReadCountsUnique"]
    counts = {}
    for j, t in enumerate(tables):
        for k, c in t.items():
            counts.setdefault(k, [0] * len(tables))[j] = c
    rows = [(k, 0, [""] * 10, counts[k]) for k in sorted(counts)]
    assert po.table_csv(rows, meta["samples"], True) == open(os.path.join(d, "complete_set.csv")).read()


SAM_FILES = {0: "miRge3_miRNA.sam", 8: "miRge3_miRNA.sam", 1: "miRge3_hairpin_miRNA.sam", 4: "miRge3_snorna.sam",
             5: "miRge3_rrna.sam", 6: "miRge3_ncrna_others.sam", 7: "miRge3_mrna.sam", 2: "miRge3_tRNA.sam",
             3: "miRge3_pre_tRNA.sam"}  # manifoldAlign.py:20-45


def test_per_round_sam_files_match_reference(oracle_run):
    """-bam / -trf: which records the reference appends to which file, in which order (the record text itself is
    the stand-in bowtie's, i.e. the oracle's sam_line)."""
    counts, annot, libs, *_ = oracle_run
    files = {}
    for rnd in range(9):  # round 9 (spike-in) is never written (manifoldAlign.py:57-62)
        lib, pol = libs[po.ROUND_LIBS[rnd]], po.ROUND_POLICIES[rnd]
        for seq in sorted(counts):
            a = annot.get(seq)
            if a is None or a[0] != rnd:
                continue
            q = po.round_query(seq, rnd)
            hs = po.hits(q, lib, pol)
            best = min(h[0] for h in hs)
            hs = sorted((h[0], h[1], h[2]) for h in hs if h[0] == best)
            if rnd not in (2, 3):
                hs = hs[:1]
            for mm, r, off in reversed(hs):
                files.setdefault(SAM_FILES[rnd], []).append(po.sam_line(seq, q, lib.names[r], lib.seqs[r], off, pol))
    assert sorted(files) == sorted(set(SAM_FILES.values()))
    for name, lines in files.items():
        assert "\n".join(lines) + "\n" == golden(os.path.join("sam", name)), name
