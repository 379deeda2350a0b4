"""Host-side logic of report.py / manifoldAlign.py that needs no GPU, against the reference-written golden files:
the canonical-ratio filter + name merging + report assembly (fed with sums computed from the golden mapped.csv on
the CPU), and the SAM tag formatting against the oracle's."""
import csv
import os

import numpy as np

from oracle import pyoracle as po
from tests.test_reference_golden import CASE, DB, ORG, SAMPLES, golden, load_case_libraries


def test_build_report_reproduces_reference_files(tmp_path):
    from mirge_b200 import report as RP

    libs = load_case_libraries()
    mir_names = libs["mirna"].names
    idx = {n: i for i, n in enumerate(mir_names)}
    rows = list(csv.DictReader(golden("mapped.csv").splitlines()))
    S = len(SAMPLES)
    sums = [(np.zeros(10, dtype=np.int64), np.zeros(len(mir_names), dtype=np.int64), np.zeros(len(mir_names), dtype=np.int64))
            for _ in range(S)]
    has_exact = np.zeros(len(mir_names), dtype=bool)
    for r in rows:
        for rnd, col in enumerate(po.ROUND_COLUMNS):
            if r[col]:
                for j, s in enumerate(SAMPLES):
                    c = int(r[s])
                    sums[j][0][rnd] += c
                    if rnd == 0:
                        sums[j][1][idx[r[col]]] += c
                    elif rnd == 8:
                        sums[j][2][idx[r[col]]] += c
                if rnd == 0:
                    has_exact[idx[r[col]]] = True
    rep = list(csv.DictReader(golden("annotation.report.csv").splitlines()))
    src = {r["Sample name(s)"]: int(r["Total Input Reads"]) for r in rep}
    trc = {r["Sample name(s)"]: int(r["Trimmed Reads (all)"]) for r in rep}
    tru = {r["Sample name(s)"]: int(r["Trimmed Reads (unique)"]) for r in rep}
    member, merged = RP.read_merges(os.path.join(CASE, "lib", ORG, "annotation.Libs", "%s_merges_%s.csv" % (ORG, DB)))
    summary, counts, rpm = RP.build_report(SAMPLES, sums, has_exact, mir_names, member, merged, src, trc, tru, 0.1, True)
    for df, name in ((summary, "annotation.report.csv"), (counts, "miR.Counts.csv"), (rpm, "miR.RPM.csv")):
        df.to_csv(tmp_path / name)
        assert (tmp_path / name).read_text() == golden(name), name
    # a missing merges file means no merging (summary.py:715-716)
    assert RP.read_merges(str(tmp_path / "absent.csv")) == ({}, [])


def test_canonical_filter_matches_the_oracle():
    from mirge_b200 import report as RP

    rng = np.random.default_rng(1)
    can = rng.integers(0, 40, 500)
    iso = rng.integers(0, 400, 500) * (rng.random(500) < 0.7)
    for thr in (0.1, 0.5, 2.0):
        got = RP.canonical_filter(can, iso, thr)
        exp = [po.canonical_filter(int(c), int(i), thr) for c, i in zip(can, iso)]
        assert got.tolist() == exp


def test_sam_tags_match_the_oracle():
    from mirge_b200 import manifoldAlign as MA

    rng = np.random.default_rng(2)
    B = np.array(list("ACGT"))
    for _ in range(300):
        L = int(rng.integers(16, 60))
        ref = "".join(rng.choice(B, L))
        q = list(ref)
        for p in rng.choice(L, int(rng.integers(0, 4)), replace=False):
            q[p] = "N" if rng.random() < 0.2 else str(rng.choice(B))
        q = "".join(q)
        seed = L if rng.random() < 0.5 else min(28, L)
        md, pos = po.md_string(q, ref)
        xa, md2, nm = MA._md_tag(q, ref, seed)
        assert (md2, nm, xa) == (md, len(pos), sum(1 for p in pos if p < seed))
    assert MA.round_query_text("ACGTACGTTTTT", po.ROUND_POLICIES[3]) == "ACGTACG"
    assert MA.round_query_text("ACGTACGTAC", po.ROUND_POLICIES[8]) == "CGTACGT"
