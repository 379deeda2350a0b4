"""CPU test of mirge_b200.launch: the reference's unchanged command line (mirge/__main__.py:main) imported with the
three hot-path modules substituted.  Needs the read-only reference checkout (skipped on the GPU box); runs in a child
process because it rewires sys.modules.  Packages this container lacks and the hot path never touches (Bio,
matplotlib: novel-miRNA / tRF / BAM code) are stubbed in the child only."""
import os
import subprocess
import sys
import textwrap

import pytest

from tests.golden import make_reference_golden as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent(
    r"""
    import sys, types
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, %(ref)r)
    def stub(name, **attrs):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m
    for missing in ("Bio", "matplotlib"):
        try:
            __import__(missing)
        except ImportError:
            if missing == "Bio":
                b = stub("Bio"); b.Seq = stub("Bio.Seq", Seq=object); b.SeqIO = stub("Bio.SeqIO"); b.pairwise2 = stub("Bio.pairwise2")
            else:
                stub("mirge.libs.novel_mir", predict_nmir=lambda *a, **k: None)  # the only importer of matplotlib
    import mirge_b200
    from mirge_b200 import launch, digest, manifoldAlign
    launch.install()
    import mirge.__main__ as M
    assert M.baking is digest.baking and M.bwtAlign is manifoldAlign.bwtAlign, "hot-path entry points not substituted"
    assert M.check_dependencies is launch._check_dependencies
    import mirge.libs.miRgeEssential as E
    assert E.validate_files.__module__ == "mirge.libs.miRgeEssential"  # the reference's own code stays
    assert "cutadapt.modifiers" not in sys.modules and "dnaio" not in sys.modules  # nothing of the CPU path was imported
    mode = sys.argv[1]
    if mode == "help":
        launch.run(["-h"])
    elif mode == "run":
        launch.run(sys.argv[2:])
    elif mode == "mEC":
        M.bakingEC(None, [], [], ".")
    """
)


def child(tmp_path, *argv):
    script = tmp_path / "child.py"
    script.write_text(CHILD % {"root": ROOT, "ref": str(G.REFERENCE)})
    return subprocess.run([sys.executable, str(script)] + list(argv), capture_output=True, text=True, cwd=str(tmp_path), timeout=600)


@pytest.fixture(autouse=True)
def need_reference():
    if not G.REFERENCE.exists():
        pytest.skip("reference checkout not present (GPU box)")


def test_reference_cli_parses_with_the_substituted_modules(tmp_path):
    p = child(tmp_path, "help")
    assert p.returncode == 0, p.stderr[-2000:]
    assert "usage:" in p.stdout.lower() and "-lib" in p.stdout  # the reference's own parser (mirge/libs/parse.py)


def test_unsupported_error_correction_stops_with_a_message(tmp_path):
    p = child(tmp_path, "mEC")
    assert p.returncode != 0 and "-mEC" in p.stderr


def test_run_reaches_the_b200_probe_and_fails_loudly_without_a_device(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-box behaviour")
    fq = tmp_path / "s1.fastq"
    fq.write_text("@r1\nACGTACGTACGTACGTACGTAC\n+\nIIIIIIIIIIIIIIIIIIIIII\n")
    lib = tmp_path / "lib"
    (lib / "human" / "index.Libs").mkdir(parents=True)
    p = child(tmp_path, "run", "-s", str(fq), "-lib", str(lib), "-on", "human", "-db", "miRBase", "-o", str(tmp_path), "-a", "illumina")
    out = p.stdout + p.stderr
    assert "no CUDA device" in out, out[-2000:]          # mirge_b200.essential.check_dependencies spoke ...
    assert "bowtie error" not in out                       # ... not the reference's probe for the bowtie binary
