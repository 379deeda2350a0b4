"""CPU tests of mirge_b200.essential: the replacement of the reference's dependency probe
(mirge/libs/miRgeEssential.py:6-98) and the ``bowtie-inspect`` shim that summarize() / bamFmt shell out to
(summary.py:776-788, :812-815, :1164-1166; bamFmt.py:10-12)."""
import argparse
import gzip
import os
import subprocess

import numpy as np
import pytest

import mirge_b200  # noqa: F401
from mirge_b200 import essential
from tests.util import write_ebwt


def make_lib(rng, n=30):
    B = np.array(list("ACGT"))
    seqs = ["".join(rng.choice(B, int(rng.integers(18, 260)))) for _ in range(n)]
    names = ["hsa-miR-%d chr1 segs:1-9,10-%d cds:+:5-9" % (i, len(s)) if i % 4 == 0 else "hsa-miR-%d" % i for i, s in enumerate(seqs)]
    return names, seqs


def run_shim(shim, *argv):
    p = subprocess.run([shim] + list(argv), capture_output=True, text=True)
    return p.returncode, p.stdout, p.stderr


def test_check_dependencies_without_a_device_exits_like_a_missing_tool(tmp_path, capsys):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-box behaviour")
    args = argparse.Namespace(quiet=False)
    log = tmp_path / "run.log"
    with pytest.raises(SystemExit):
        essential.check_dependencies(args, log)
    assert "no CUDA device" in log.read_text()
    assert "no CUDA device" in capsys.readouterr().out
    assert not hasattr(args, "cutadaptVersion")


def test_check_dependencies_sets_the_attributes_the_reference_reads(tmp_path, monkeypatch, capsys):
    import torch

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "get_device_name", lambda i: "NVIDIA B200")
    args = argparse.Namespace(quiet=True)
    log = tmp_path / "run.log"
    essential.check_dependencies(args, log)
    assert args.bowtieVersion == "True"  # novel_mir.py:319, mirge2_tRF_a2i.py:1057 compare with the string
    assert int(args.cutadaptVersion[0]) >= 3  # digest.py:111-112 reads element 0
    text = log.read_text()
    assert "bowtie version: " in text and "cutadapt version: " in text and "NVIDIA B200" in text
    assert capsys.readouterr().out == ""  # quiet


@pytest.mark.parametrize("source", ["fasta", "fasta_gz_in_fasta_libs", "ebwt"])
def test_inspect_shim_answers_the_reference_calls(tmp_path, source):
    rng = np.random.default_rng(11)
    names, seqs = make_lib(rng)
    idx_dir = tmp_path / "human" / "index.Libs"
    idx_dir.mkdir(parents=True)
    base = str(idx_dir / "human_mirna_miRBase")
    text = "".join(">%s\n%s\n" % (n, s) for n, s in zip(names, seqs))
    if source == "fasta":
        open(base + ".fa", "w").write(text)
    elif source == "fasta_gz_in_fasta_libs":
        (tmp_path / "human" / "fasta.Libs").mkdir()
        with gzip.open(str(tmp_path / "human" / "fasta.Libs" / "human_mirna_miRBase.fasta.gz"), "wt") as f:
            f.write(text)
    else:
        write_ebwt(base, names, seqs)
    shim = essential.write_inspect_shim(str(tmp_path / "bin"))
    assert os.access(shim, os.X_OK)
    # summary.py:776-788: one name line per reference; lines with "segs:" are cut at the first blank by the caller
    rc, out, err = run_shim(shim, "-n", base)
    assert rc == 0, err
    assert out.strip().split("\n") == names
    merged = [r.split(" ")[0] if "segs:" in r else r for r in out.strip().split("\n")]
    assert merged == [n.split(" ")[0] for n in names]
    # summary.py:812-815 / :1164-1166: "-a 20000 -e" = one sequence line per reference
    rc, out, err = run_shim(shim, "-a", "20000", "-e", base)
    assert rc == 0, err
    lines = out.strip().split("\n")
    assert lines[0::2] == [">" + n for n in names] and lines[1::2] == seqs
    # default line width of bowtie-inspect
    rc, out, _ = run_shim(shim, base)
    got = {}
    for ln in out.strip().split("\n"):
        if ln.startswith(">"):
            cur = ln[1:]
            got[cur] = ""
        else:
            assert len(ln) <= 60
            got[cur] += ln
    assert got == dict(zip(names, seqs))


def test_inspect_shim_fails_loudly(tmp_path):
    shim = essential.write_inspect_shim(str(tmp_path / "bin"))
    rc, out, err = run_shim(shim, "-n", str(tmp_path / "absent_index"))
    assert rc != 0 and "missing" in err and out == ""
    rc, _, err = run_shim(shim, "--threads", "4", "x")
    assert rc != 0 and "unsupported option" in err
    rc, _, err = run_shim(shim, "-n")
    assert rc != 0


def test_inspect_main_in_process(tmp_path):
    import io

    base = str(tmp_path / "lib")
    open(base + ".fa", "w").write(">a desc\nACGT\nAC\n>b\nTTTT\n")
    buf = io.StringIO()
    assert essential.inspect_main(["-a", "4", base], out=buf) == 0
    assert buf.getvalue() == ">a desc\nACGT\nAC\n>b\nTTTT\n"
    assert essential.index_names(base) == ["a desc", "b"]
    assert essential.index_entries(base) == (["a desc", "b"], [b"ACGTAC", b"TTTT"])


def test_unmodified_reference_summarize_runs_on_the_shim(tmp_path):
    """The reference's own summarize() (mirge/libs/summary.py, imported from the read-only checkout when it is present)
    with args.bowtie_path pointing at the shim: the report files must be the golden ones, which the generator produced
    with its private stand-in -- i.e. the product shim is a drop-in for ``bowtie-inspect -n`` there."""
    from pathlib import Path

    import pandas as pd

    from tests.golden import make_reference_golden as G
    from tests.test_reference_golden import CASE, DB, SAMPLES, golden

    if not G.REFERENCE.exists():
        pytest.skip("reference checkout not present (GPU box)")
    from tests.golden import standins

    standins.install(G.REFERENCE)
    from mirge.libs.summary import summarize  # the reference's code, unmodified

    bindir = tmp_path / "bin"
    essential.write_inspect_shim(str(bindir))
    work = tmp_path / "work"
    work.mkdir()
    args = G.reference_args(Path(CASE) / "lib", bindir, quality_cutoff="20")
    mapped = pd.read_csv(os.path.join(CASE, "mapped.csv"), index_col="Sequence", keep_default_na=False)
    rep = pd.read_csv(os.path.join(CASE, "annotation.report.csv"))
    src = dict(zip(rep["Sample name(s)"], rep["Total Input Reads"].astype(int)))
    trc = dict(zip(rep["Sample name(s)"], rep["Trimmed Reads (all)"].astype(int)))
    tru = dict(zip(rep["Sample name(s)"], rep["Trimmed Reads (unique)"].astype(int)))
    cwd = os.getcwd()
    try:
        summarize(args, work, DB, list(SAMPLES), mapped, src, trc, tru)
    finally:
        os.chdir(cwd)
    for name in ("annotation.report.csv", "miR.Counts.csv", "miR.RPM.csv"):
        assert (work / name).read_text() == golden(name), name


def test_downstream_tools_are_probed_before_the_digest(tmp_path, monkeypatch):
    """launch._check_dependencies: the tools the reference's downstream modules shell out to (novel_mir.py:224-225,318-323,
    mirge2_tRF_a2i.py:1056,1291, bamFmt.py:116,174) are looked for up front, and the bowtie-inspect shim directory does not
    hide a real bowtie / bowtie-build behind args.bowtie_path."""
    import argparse
    import stat

    from mirge_b200 import launch

    a = argparse.Namespace(novel_miRNA=True, AtoI=False, tRNA_frag=True, bam_out=True, bowtie_path=None, samtools_path=None, RNAfold_path=None)
    need = launch.downstream_tools(a)
    assert {(n, o) for _, n, o in need} == {("bowtie", "-nmir"), ("bowtie-build", "-nmir"), ("samtools", "-nmir"), ("RNAfold", "-nmir"),
                                          ("bowtie", "-trf"), ("samtools", "-bam")}
    assert launch.downstream_tools(argparse.Namespace()) == []
    # a directory with fake tools: found through -pbwt, not through PATH
    bindir = tmp_path / "bin"
    bindir.mkdir()
    for name in ("bowtie", "bowtie-build"):
        p = bindir / name
        p.write_text("#!/bin/sh\nexit 0\n")
        p.chmod(p.stat().st_mode | stat.S_IXUSR)
    monkeypatch.setenv("PATH", str(tmp_path / "nothing"))
    a.bowtie_path = str(bindir)
    assert launch._tool(a, "bowtie_path", "bowtie") == str(bindir / "bowtie")
    assert launch._tool(a, "bowtie_path", "bowtie-inspect") is None
    assert launch._tool(a, "samtools_path", "samtools") is None
    # the full probe: samtools / RNAfold missing -> early exit naming them and the options
    monkeypatch.setattr("mirge_b200.essential.check_dependencies", lambda args, log: None)
    with pytest.raises(SystemExit) as e:
        launch._check_dependencies(a, str(tmp_path / "run.log"))
    assert "samtools (needed by -bam)" in str(e.value) and "RNAfold (needed by -nmir)" in str(e.value) and "bowtie (" not in str(e.value)
    # nothing missing: the shim directory becomes args.bowtie_path and carries links to the real bowtie / bowtie-build
    b = argparse.Namespace(novel_miRNA=False, AtoI=True, tRNA_frag=False, bam_out=False, bowtie_path=str(bindir), samtools_path=None,
                           RNAfold_path=None)
    launch._check_dependencies(b, str(tmp_path / "run.log"))
    assert b.bowtie_path.endswith(".mirge_b200_bin")
    for name in ("bowtie-inspect", "bowtie", "bowtie-build"):
        assert os.access(os.path.join(b.bowtie_path, name), os.X_OK)
    assert os.path.realpath(os.path.join(b.bowtie_path, "bowtie")) == os.path.realpath(str(bindir / "bowtie"))
