"""Linked adapters ``-g "ADAPTER5...ADAPTER3"`` (the reference documents them: docs/source/quick_start.md:208-220; the
semantics are cutadapt's LinkedAdapter / LinkedMatch, a third-party dependency absent from /root/reference and restated
in oracle/pyoracle.py::match_linked): known answers worked out by hand, on the Python oracle and on the C oracle through
the flattened parameters the kernels receive.  CPU only."""
import numpy as np

import mirge_b200  # noqa: F401
from mirge_b200 import params as P
from oracle import coracle
from oracle import pyoracle as po

FIVE, THREE = "TTAGGC", "TGGAATTCTCGGGTGCCAAGGAACTCCAGT"
INS = "TAGCTTATCAGACTGATGTTGA"


def pair(**kw):
    return po.LinkedAdapter(po.Adapter("front", FIVE, **kw), po.Adapter("back", THREE, **kw), True, True)


def test_linked_match_known_answers():
    ad = pair()
    # both halves: everything up to the end of the 5' adapter and from the start of the 3' adapter goes
    r = "AC" + FIVE + INS + THREE + "CAC"
    assert po.match_linked(ad, r) == (0, 0, 8, 8 + len(INS), len(FIVE) + len(THREE), 0)
    # -g requires both halves (cutadapt parser: front_required = back_required = True)
    assert po.match_linked(ad, INS + THREE) is None
    assert po.match_linked(ad, "AC" + FIVE + INS) is None
    # a partial 5' adapter at the read start (5' adapters may be cut off at the front) and a partial 3' adapter at the end
    r = FIVE[2:] + INS + THREE[:12]
    assert po.match_linked(ad, r) == (0, 0, 4, 4 + len(INS), 4 + 12, 0)
    # the 3' half is searched only in what the 5' match leaves: a 3' adapter copy in front of the 5' adapter is not seen
    r = THREE[:15] + FIVE + INS
    assert po.match_linked(ad, r) is None
    # one substitution in the 3' half: errors add up over the halves
    t3 = THREE[:10] + ("A" if THREE[10] != "A" else "C") + THREE[11:]
    mt = po.match_linked(ad, FIVE + INS + t3)
    assert mt[2:4] == (6, 6 + len(INS)) and mt[5] == 1 and mt[4] == len(FIVE) + len(THREE) - 1


def test_linked_pipeline_python_and_c_oracle_agree_with_the_hand_result():
    cfg = P.TrimConfig(adapters=[("front", FIVE + "..." + THREE)], quality_cutoff=None, minimum_length=16)
    reads = ["AC" + FIVE + INS + THREE + "CAC", INS + THREE, "AC" + FIVE + INS, FIVE[2:] + INS + THREE[:12]]
    fq = "".join("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)) for i, s in enumerate(reads)).encode()
    cp = P.build_trim_params(cfg)
    n, tab = coracle.digest_collapse(np.frombuffer(fq, dtype=np.uint8), cp)
    # reads 0 and 3 are cut to the insert; reads 1 and 2 stay whole (no pair found), HEAD counting: one modifier -> one count
    exp = {INS: 2, reads[1]: 1, reads[2]: 1}
    assert n == 4 and tab.to_dict() == exp
    pp = po.TrimParams(adapters=[pair()], quality_cutoff=None, minimum_length=16)
    n_py, got = po.digest_chunk(fq, pp)
    assert n_py == 4 and got == exp


def test_min_overlap_is_lowered_to_the_adapter_length():
    """cutadapt's Adapter lowers min_overlap to len(sequence); without it a 6-nt adapter under -O 7 could only be found
    by match_to's exact-find shortcut and never by the alignment -- the two must agree (and the kernels only align)."""
    ad = po.Adapter("front", "CATGTC", 0.34, 7)
    assert ad.min_overlap == 6
    cp = P.build_trim_params(P.TrimConfig(adapters=[("front", "CATGTC")], overlap=7))
    assert cp.adapters[0].min_overlap == 6
    rng = np.random.default_rng(4)
    for _ in range(300):
        read = "".join(rng.choice(list("ACGT"), int(rng.integers(0, 30)))) + "CATGTC" + "".join(rng.choice(list("ACGT"), int(rng.integers(0, 30))))
        assert po.match_to(ad, read) == po.locate(ad, read)
