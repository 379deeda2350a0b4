"""GPU twin of tests/test_reference_golden.py: the product entry points (baking -> bwtAlign -> report) must
reproduce, byte for byte, the files the UNMODIFIED reference wrote for the same inputs
(tests/golden/ref_case1: mapped.csv, unmapped.csv, annotation.report.csv, miR.Counts.csv)."""
import os

import numpy as np
import pytest

from tests.test_gpu_annotate import make_args
from tests.test_reference_golden import CASE, DB, ORG, SAMPLES, golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    from mirge_b200 import device as D

    return D.Device(0)


def test_pipeline_reproduces_reference_files(dev, tmp_path):
    from mirge_b200 import digest as DG
    from mirge_b200 import manifoldAlign as MA
    from mirge_b200 import report as RP

    args = make_args(libraries_path=os.path.join(CASE, "lib"), organism_name=ORG, spikeIn=True, quality_cutoff="20",
                     crThreshold="0.1")
    files = [os.path.join(CASE, s + ".fastq.gz") for s in SAMPLES]
    df, src, trc, tru = DG.baking(args, files, SAMPLES, str(tmp_path), device=dev, batch_bytes=150_000)
    out = MA.bwtAlign(args, df, str(tmp_path), DB, device=dev)
    RP.write_tables(out, str(tmp_path))
    summary, mir, rpm = RP.annotation_report(args, str(tmp_path), DB, SAMPLES, out, src, trc, tru, device=dev)
    for name in ("mapped.csv", "unmapped.csv", "annotation.report.csv", "miR.Counts.csv", "miR.RPM.csv"):
        assert (tmp_path / name).read_text() == golden(name), name
    assert summary.loc["sampleA", "Total Input Reads"] == 2500


def test_device_resident_sums_match_dataframe_path(dev, tmp_path):
    """SampleSums fed straight from the device arrays of the pipeline (no DataFrame) gives the same counters."""
    import gzip

    from mirge_b200 import device as D
    from mirge_b200 import libraries as LB
    from mirge_b200 import manifoldAlign as MA
    from mirge_b200 import params as P
    from mirge_b200 import report as RP

    args = make_args(libraries_path=os.path.join(CASE, "lib"), organism_name=ORG, spikeIn=True, quality_cutoff="20")
    libs = MA.load_libraries(args, DB, dev)
    eng = D.DigestEngine(dev, P.TrimConfig.from_args(args, "head"))
    table = D.CollapseTable(dev, min_keys=1 << 10)
    per_sample = []
    for s in SAMPLES:
        with gzip.open(os.path.join(CASE, s + ".fastq.gz"), "rb") as f:
            buf = torch.frombuffer(bytearray(f.read()), dtype=torch.uint8).to(dev.tdev)
        eng.digest_device(buf, table)
        per_sample.append(table.drain())
    annot, hit = MA.annotate_keys(dev, libs, MA.KeySet.from_table(table), True)
    want = [l.split(",") for l in golden("annotation.report.csv").splitlines()]
    col = {n: i for i, n in enumerate(want[0])}
    for j, (ids, cnt) in enumerate(per_sample):
        ss = RP.SampleSums(dev, libs["mirna"].n_refs)
        ss.add(annot, hit, ids, cnt)
        rs, can, iso = ss.host()
        row = want[1 + j]
        assert int(rs[0] + rs[8]) == int(row[col["All miRNA Reads"]])
        for name, rnd in RP._ROUND_OF.items():
            assert int(rs[rnd]) == int(row[col[name]]), name
        assert int(can.sum()) == int(rs[0]) and int(iso.sum()) == int(rs[8])


from tests.test_reference_golden import DIGEST_CASES, load_digest_case, sorted_lines, tcf_pairs  # noqa: E402


@pytest.mark.parametrize("name", DIGEST_CASES)
def test_baking_matches_reference_baking(dev, tmp_path, name):
    """Product baking() vs the files the reference's own baking() wrote (UMI flanks with and without -udd,
    qiagen UMIs, NextSeq + two-sided quality cut-offs + -NX + -u cuts + times=2; side files included)."""
    from mirge_b200 import digest as DG

    d, meta, args, files = load_digest_case(name)
    df, src, trc, tru = DG.baking(args, files, meta["samples"], str(tmp_path), device=dev, batch_bytes=70_000)
    assert src == meta["sampleReadCounts"] and trc == meta["trimmedReadCounts"] and tru == meta["trimmedReadCountsUnique"]
    df.to_csv(tmp_path / "complete_set.csv")
    assert (tmp_path / "complete_set.csv").read_text() == open(os.path.join(d, "complete_set.csv")).read()
    assert {c: str(t) for c, t in df.dtypes.items() if str(t) == "int64"} == {c: t for c, t in meta["dtypes"].items() if t == "int64"}
    for s in meta["samples"]:
        umi_csv = os.path.join(d, s + "_umiCounts.csv")
        if os.path.exists(umi_csv):
            assert sorted_lines(str(tmp_path / (s + "_umiCounts.csv"))) == sorted_lines(umi_csv)
        tcf = os.path.join(d, s + ".trim.collapse.fa")
        if os.path.exists(tcf):
            assert tcf_pairs(str(tmp_path / (s + ".trim.collapse.fa"))) == tcf_pairs(tcf)


def test_per_round_sam_files_match_reference(dev, tmp_path):
    """bwtAlign with -bam and -trf writes the per-round SAM files the reference wrote (all best-stratum hits for
    the two tRNA rounds, canonical pick last), and the table is the same as without the flags."""
    from mirge_b200 import digest as DG
    from mirge_b200 import manifoldAlign as MA

    args = make_args(libraries_path=os.path.join(CASE, "lib"), organism_name=ORG, spikeIn=True, quality_cutoff="20",
                     bam_out=True, tRNA_frag=True)
    files = [os.path.join(CASE, s + ".fastq.gz") for s in SAMPLES]
    df, *_ = DG.baking(args, files, SAMPLES, str(tmp_path), device=dev)
    out = MA.bwtAlign(args, df, str(tmp_path), DB, device=dev)
    out.to_csv(tmp_path / "all.csv")
    sam_dir = os.path.join(CASE, "sam")
    for name in sorted(os.listdir(sam_dir)):
        assert (tmp_path / name).read_text() == open(os.path.join(sam_dir, name)).read(), name
    assert int((out.annotFlag == 1).sum()) == len(golden("mapped.csv").splitlines()) - 1
