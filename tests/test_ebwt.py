"""Loader for bowtie 1 index files (mirge_b200/ebwt.py) against indexes written by tests.util.write_ebwt, which
follows the same published layout (no real bowtie-build exists here: this checks self-consistency and the
loader's refusal to guess, not the layout itself)."""
import numpy as np
import pytest

from tests.util import write_ebwt


def test_round_trip_with_ambiguous_bases_and_descriptions(tmp_path):
    from mirge_b200 import ebwt

    rng = np.random.default_rng(5)
    B = np.array(list("ACGT"))
    seqs = ["".join(rng.choice(B, int(rng.integers(18, 400)))) for _ in range(40)]
    seqs[3] = seqs[3][:10] + "NNN" + seqs[3][13:]          # internal ambiguity is kept as N
    seqs[7] = "NN" + seqs[7]                                # leading, too
    seqs[9] = seqs[9] + "NNNN"                              # trailing ambiguity is not recorded by the index
    names = ["ref%d chr1 segs:1-9,10-%d cds:+:5-9" % (i, len(s)) if i % 5 == 0 else "ref%d" % i for i, s in enumerate(seqs)]
    base = str(tmp_path / "human_mirna_miRBase")
    write_ebwt(base, names, seqs)
    got_names, got = ebwt.decode_index(base)
    assert got_names == ["ref%d" % i for i in range(40)]
    assert ebwt.decode_names(base) == names
    exp = [s.encode() for s in seqs]
    exp[9] = seqs[9][:-4].encode()
    assert got == exp


def test_name_block_found_when_the_header_arithmetic_does_not_fit(tmp_path):
    from mirge_b200 import ebwt

    base = str(tmp_path / "lib")
    write_ebwt(base, ["a", "b desc"], ["ACGTACGTAC", "TTTTGGGGCCCCAAAA"], bwt_pad=52)
    assert ebwt.decode_index(base) == (["a", "b"], [b"ACGTACGTAC", b"TTTTGGGGCCCCAAAA"])


def test_refuses_inconsistent_indexes(tmp_path):
    from mirge_b200 import ebwt
    from mirge_b200.device import MirgeError

    base = str(tmp_path / "lib")
    write_ebwt(base, ["a", "b"], ["ACGTACGTAC", "TTTTGGGG"])
    raw = bytearray(open(base + ".1.ebwt", "rb").read())
    open(base + ".1.ebwt", "wb").write(bytes(raw[:-1]))                      # closing NUL lost
    with pytest.raises(MirgeError):
        ebwt.decode_index(base)
    write_ebwt(base, ["a", "b", "c"], ["ACGTACGTAC", "TTTTGGGG"])              # one name too many
    with pytest.raises(MirgeError):
        ebwt.decode_index(base)
    write_ebwt(base, ["a", "b"], ["ACGTACGTAC", "TTTTGGGG"])
    open(base + ".4.ebwt", "wb").write(b"\x00")                               # bases missing
    with pytest.raises(MirgeError):
        ebwt.decode_index(base)
    with pytest.raises(MirgeError):
        ebwt.decode_index(str(tmp_path / "absent"))
