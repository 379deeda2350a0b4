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
