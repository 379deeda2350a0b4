"""CPU tests of mirge_b200.params: the adapter specification subset, the flattened trim parameters (the product's
``stipulate()``), and -- when the read-only reference checkout is present -- the modifier pipeline held against the
list the reference's own ``stipulate(args)`` (mirge/libs/digest.py:59-101) builds for the same arguments."""
import subprocess
import sys
import textwrap

import pytest

import mirge_b200  # noqa: F401
from mirge_b200 import abi
from mirge_b200 import params as P
from tests.golden import make_reference_golden as G

ILL_BACK = "TGGAATTCTCGGGTGCCAAGGAACTCCAG"
ILL_FRONT = "GTTCAGAGTTCTACAGTCCGACGATC"


def test_adapter_specification_subset():
    assert P.parse_adapter_spec("back", "illumina").sequence == ILL_BACK       # mirge/__main__.py:66-86
    assert P.parse_adapter_spec("front", "Illumina").sequence == ILL_FRONT
    assert P.parse_adapter_spec("back", "myname=acgu").sequence == "ACGT"      # name=, upper-casing, U -> T
    assert P.parse_adapter_spec("back", " ACGTN ").sequence == "ACGTN"
    for kind, spec in (("back", "file:adapters.fa"), ("back", "ACGT...TTTT"), ("back", "ACGT$"), ("front", "^ACGT"),
                       ("back", "ACGT;max_error_rate=0.2"), ("back", "ACGTX"), ("front", "XACGT"), ("back", ""),
                       ("back", "A" * (abi.MAX_ADAPTER_LEN + 1)), ("back", "ACGT-ACGT"), ("anywhere", "ACGT")):
        with pytest.raises(P.UnsupportedAdapterSpec):
            P.parse_adapter_spec(kind, spec)


def test_linked_adapter_specification():
    """-g "ADAPTER5...ADAPTER3" (docs/source/quick_start.md:208-220): a 5' half pointing at its 3' half in the flat adapter
    list; the forms that anchor a half, and the alias inside a linked specification, stay rejected."""
    sp = P.parse_adapter_spec("front", "TTAGGC...TGGAATTCTCGGGTGCCAAGGAACTCCAGT")
    assert (sp.where, sp.sequence, sp.sequence2) == ("linked", "TTAGGC", "TGGAATTCTCGGGTGCCAAGGAACTCCAGT")
    p = P.build_trim_params(P.TrimConfig(adapters=[("back", "ACGTACGTAC"), ("front", "name=TTAGGC...ACGTTGCA")]))
    assert p.n_adapters == 3
    assert [p.adapters[i].where for i in range(3)] == [0, 1, 0]
    assert [p.adapters[i].link for i in range(3)] == [0, 3, abi.LINK_BACK_HALF]
    for kind, spec in (("front", "TTAGGC...illumina"), ("front", "^TTAGGC...ACGT"), ("front", "TTAGGC...ACGT$"), ("front", "A...C...G"),
                       ("front", "...ACGT"), ("back", "TTAGGC...ACGT")):
        with pytest.raises(P.UnsupportedAdapterSpec):
            P.parse_adapter_spec(kind, spec)
    with pytest.raises(P.UnsupportedAdapterSpec):
        P.build_trim_params(P.TrimConfig(adapters=[("front", "TTAGGC...AACTGTAGGCACCATCAAT")], qiagenumi=True, uniq_mol_ids="0,12"))


def test_parse_cutoffs_doctests():  # digest.py:19-35
    assert P.parse_cutoffs("5") == [0, 5] or tuple(P.parse_cutoffs("5")) == (0, 5)
    assert tuple(P.parse_cutoffs("6,7")) == (6, 7)
    with pytest.raises(Exception):
        P.parse_cutoffs("1,2,3")


def kinds(p):
    return [p.mod_kind[i] for i in range(p.n_mods)]


def test_modifier_order_and_parameters():
    cfg = P.TrimConfig(adapters=[("back", "illumina")], nextseq_trim=20, quality_cutoff="5,15", quality_base=64, trim_n=True, cut=[2, -3],
                       times=2, minimum_length=14, error_rate=0.12, overlap=3)
    p = P.build_trim_params(cfg)
    assert kinds(p) == [abi.MOD_NEXTSEQ, abi.MOD_QUALITY, abi.MOD_ADAPTER, abi.MOD_NEND, abi.MOD_CUT, abi.MOD_CUT]  # digest.py:87-99
    assert (p.mod_a[0], p.mod_b[0]) == (20, 64)
    assert (p.mod_a[1], p.mod_b[1], p.mod_c[1]) == (5, 15, 64)
    assert (p.mod_a[4], p.mod_a[5]) == (2, -3)
    assert p.times == 2 and p.min_len == 14 and p.n_adapters == 1
    assert P.trim_slots(p) == 6                                               # HEAD: one emission slot per modifier
    assert P.trim_slots(P.build_trim_params(P.TrimConfig(adapters=[("back", "illumina")], count_mode="release"))) == 1
    a = p.adapters[0]
    assert a.m == 29 and a.k == int(0.12 * 29) and a.min_overlap == 3 and a.indel_cost == 1 and a.wildcard_ref == 0
    # cutadapt compares cost <= length * rate in double: the integer table must be its floor for every length
    for rate in (0.0, 0.05, 0.1, 0.12, 0.2, 0.34, 0.5):
        ad = P.build_adapter(P.parse_adapter_spec("back", ILL_BACK), P.TrimConfig(adapters=[("back", ILL_BACK)], error_rate=rate))
        for L in range(abi.MAX_ADAPTER_LEN + 1):
            assert all((c <= L * rate) == (c <= ad.max_err[L]) for c in range(0, 40)), (rate, L)
    # a cut of zero adds no modifier (digest.py:51-53); no quality cut-off, no adapter: empty pipeline
    assert kinds(P.build_trim_params(P.TrimConfig(adapters=[], quality_cutoff=None, cut=[0]))) == []


def test_wildcard_adapters_and_rejections():
    cfg = P.TrimConfig(adapters=[("back", "ACGTNNACGR")])
    a = P.build_trim_params(cfg).adapters[0]
    assert a.wildcard_ref == 1 and a.effective_length == 8 and [a.n_counts[i] for i in range(11)] == [0, 0, 0, 0, 0, 1, 2, 2, 2, 2, 2]
    with pytest.raises(P.UnsupportedAdapterSpec):
        P.build_trim_params(P.TrimConfig(adapters=[("back", "NNNN")]))                      # cutadapt: only N wildcards
    with pytest.raises(P.UnsupportedAdapterSpec):
        P.build_trim_params(P.TrimConfig(adapters=[("back", "ACGTN")], match_adapter_wildcards=False))
    with pytest.raises(RuntimeError):
        P.build_trim_params(P.TrimConfig(adapters=[("back", "ACGT")], match_read_wildcards=True))
    with pytest.raises(RuntimeError):
        P.build_trim_params(P.TrimConfig(adapters=[("back", "ACGT")], cut=[1, 2, 3]))          # digest.py:47-48
    with pytest.raises(RuntimeError):
        P.build_trim_params(P.TrimConfig(adapters=[("back", "ACGT")], cut=[1, 2]))             # digest.py:49-50
    with pytest.raises(RuntimeError):
        P.build_trim_params(P.TrimConfig(adapters=[("back", "ACGT")], qiagenumi=True))         # needs -umi
    with pytest.raises(RuntimeError):
        P.build_trim_params(P.TrimConfig(adapters=[("back", "ACGT")], action="mask"))
    q = P.build_trim_params(P.TrimConfig(adapters=[("back", "AACTGTAGGCACCATCAAT")], qiagenumi=True, uniq_mol_ids="0,12"))
    assert q.umi_mode == abi.UMI_QIAGEN and q.qia_adapter_len == 19 and (q.umi5, q.umi3) == (0, 12) and P.trim_slots(q) == 1


CHILD = textwrap.dedent(
    r"""
    import json, sys
    sys.path.insert(0, %(root)r)
    from tests.golden import make_reference_golden as G, standins
    standins.install(G.REFERENCE)
    from mirge.libs.digest import stipulate   # the reference's own code (third-party classes are the stand-ins)
    out = []
    for ov in json.loads(sys.argv[1]):
        if "adapters" in ov:
            ov["adapters"] = [tuple(a) for a in ov["adapters"]]
        mods = stipulate(G.reference_args(**ov))
        row = []
        for m in mods:
            n = type(m).__name__
            if n == "NextseqQualityTrimmer": row.append(["nextseq", m.cutoff, m.base])
            elif n == "QualityTrimmer": row.append(["quality", m.cf, m.cb, m.base])
            elif n == "AdapterCutter": row.append(["adapter", m.times, [[a.where, a.sequence] for a in m.adapters]])
            elif n == "NEndTrimmer": row.append(["nend"])
            elif n == "UnconditionalCutter": row.append(["cut", m.length])
            else: row.append([n])
        out.append(row)
    print(json.dumps(out))
    """
)


def test_pipeline_equals_the_reference_stipulate(tmp_path):
    import json
    import os

    if not G.REFERENCE.exists():
        pytest.skip("reference checkout not present (GPU box)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cases = [
        dict(),
        dict(nextseq_trim=20, quality_cutoff="20"),
        dict(nextseq_trim=15, quality_cutoff="5,15", trim_n=True, cut=[2, -3], times=2),
        dict(quality_cutoff=None, trim_n=True, cut=[-4]),
        dict(adapters=[["back", ILL_BACK], ["front", ILL_FRONT]], indels=False, cut=[0, 3]),
        dict(adapters=[], quality_cutoff="7", phred64=64),
    ]
    script = tmp_path / "child.py"
    script.write_text(CHILD % {"root": root})
    p = subprocess.run([sys.executable, str(script), json.dumps(cases)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    ref = json.loads(p.stdout.strip().splitlines()[-1])
    names = {abi.MOD_NEXTSEQ: "nextseq", abi.MOD_QUALITY: "quality", abi.MOD_ADAPTER: "adapter", abi.MOD_NEND: "nend", abi.MOD_CUT: "cut"}
    for ov, row in zip(cases, ref):
        ov = dict(ov)
        if "adapters" in ov:
            ov["adapters"] = [tuple(a) for a in ov["adapters"]]
        args = G.reference_args(**ov)
        cfg = P.TrimConfig.from_args(args)
        tp = P.build_trim_params(cfg)
        mine = []
        for i in range(tp.n_mods):
            k = names[tp.mod_kind[i]]
            if k == "nextseq":
                mine.append([k, tp.mod_a[i], tp.mod_b[i]])
            elif k == "quality":
                mine.append([k, tp.mod_a[i], tp.mod_b[i], tp.mod_c[i]])
            elif k == "adapter":
                mine.append([k, tp.times, [["back" if tp.adapters[a].where == 0 else "front",
                                            bytes(tp.adapters[a].ascii[: tp.adapters[a].m]).decode()] for a in range(tp.n_adapters)]])
            elif k == "cut":
                mine.append([k, tp.mod_a[i]])
            else:
                mine.append([k])
        assert mine == row, (ov, mine, row)
