"""Pin the oracle on every known-answer vector the reference repo holds for this path
(SURVEY.md section 4): the ten documented reads in docs/source/quick_start.md:285-315, the
parse_cutoffs doctest (digest.py:23-24), and hand-derived cutadapt edge cases (Appendix A)."""
import pytest

from oracle import pyoracle as po

LET7A = "TGAGGTAGTAGGTTGTATAGTT"
QIA_AD = "AACTGTAGGCACCATCAAT"
ILL_AD = "TGGAATTCTCGGGTGCCAAGGAACTCCAG"  # __main__.py:68 ("illumina" alias)

QIA_READS = [  # docs/source/quick_start.md:292-298
    "TGAGGTAGTAGGTTGTATAGTTAACTGTAGGCACCATCAATGTTAGACCTGCAAGATCGGAAGAGCACACGTCTG",
    "TGAGGTAGTAGGTTGTATAGTTAACTGTAGGCACCATCAATCAATGACGATTTAGATCGGAAGAGCACACGTCTG",
    "TGAGGTAGTAGGTTGTATAGTTAACTGTAGGCACCATCAATAAACAAAGATCCAGATCGGAAGAGCACACGTCTG",
    "TGAGGTAGTAGGTTGTATAGTTAACTGTAGGCACCATCAATCGCATCGCCGACAGATCGGAAGAGCACACGTCTG",
    "TGAGGTAGTAGGTTGTATAGTTAACTGTAGGCACCATCAATTTTGCCATTACTAGATCGGAAGAGCACACGTCTG",
]
QIA_UMIS = ["GTTAGACCTGCA", "CAATGACGATTT", "AAACAAAGATCC", "CGCATCGCCGAC", "TTTGCCATTACT"]

ILL_READS = [  # docs/source/quick_start.md:309-315 (read 4 has a C->A error inside the adapter)
    "TACATGAGGTAGTAGGTTGTATAGTTCCTCTGGAATTCTCGGGTGCCAAGGAACTCCAGTCACCGGAATATCTCG",
    "TACCTGAGGTAGTAGGTTGTATAGTTACTATGGAATTCTCGGGTGCCAAGGAACTCCAGTCACCGGAATATCTCG",
    "CAGGTGAGGTAGTAGGTTGTATAGTTGGTATGGAATTCTCGGGTGCCAAGGAACTCCAGTCACCGGAATATCTCG",
    "AGAATGAGGTAGTAGGTTGTATAGTTACTATGGAATTCTCGGGTGACAAGGAACTCCAGTCACCGGAATATCTCG",
    "AGGTTGAGGTAGTAGGTTGTATAGTTACTATGGAATTCTCGGGTGCCAAGGAACTCCAGTCACCGGAATATCTCG",
]
ILL_UMIS = [("TACA", "CCTC"), ("TACC", "ACTA"), ("CAGG", "GGTA"), ("AGAA", "ACTA"), ("AGGT", "ACTA")]


def fq(reads, q="I"):
    return "".join("@r%d\n%s\n+\n%s\n" % (i, r, q * len(r)) for i, r in enumerate(reads)).encode()


def test_parse_cutoffs_doc():
    assert po.parse_cutoffs("5") == [0, 5]
    assert po.parse_cutoffs("6,7") == [6, 7]


def test_qiagen_doc_reads():
    # miRge3.0 ... -a AACTGTAGGCACCATCAAT --qiagenumi -umi 0,12 -udd   (quick_start.md:285)
    p = po.TrimParams(adapters=[po.Adapter("back", QIA_AD)], umi=(0, 12), qiagenumi=True)
    for read, umi in zip(QIA_READS, QIA_UMIS):
        out = po.digest_read(read, "I" * len(read), p)
        assert [k for k, _ in out] == [LET7A + umi]
    d = po.digest_sample(fq(QIA_READS), p, umi_dedup=True)
    assert d.count == 5 and d.table == {LET7A: 5} and d.trimmed == 5
    assert sorted(d.umi_rows) == sorted((u, LET7A, 1) for u in QIA_UMIS)
    d = po.digest_sample(fq(QIA_READS + QIA_READS[:2]), p, umi_dedup=False)
    assert d.table == {LET7A: 7} and d.trimmed == 7 and sorted(d.hist) == [1, 1, 1, 2, 2]


def test_illumina_4n_doc_reads():
    # miRge3.0 ... -a illumina -umi 4,4 -udd   (quick_start.md:304)
    ad = po.Adapter("back", ILL_AD)
    for read, (u5, u3) in zip(ILL_READS, ILL_UMIS):
        mt = po.match_to(ad, read)
        assert mt is not None
        trimmed = read[: mt[2]]
        assert po.umi_parser(trimmed, 4, 4) == (LET7A, u5 + u3)
    # read 4: one substitution inside the 29-nt adapter -> cost 1, 28 matches
    assert po.match_to(ad, ILL_READS[3])[4:] == (28, 1)
    assert po.match_to(ad, ILL_READS[0])[4:] == (29, 0)
    p = po.TrimParams(adapters=[ad], umi=(4, 4), count_mode="release")
    d = po.digest_sample(fq(ILL_READS), p, umi_dedup=True)
    assert d.table == {LET7A: 5} and d.trimmed == 5
    # HEAD counts after every modifier (digest.py:354-373): the untrimmed 75-nt read is a key too
    p = po.TrimParams(adapters=[ad], umi=(4, 4), count_mode="head")
    d1 = po.digest_sample(fq(ILL_READS), p, umi_dedup=False)
    assert d1.table[LET7A] == 5 and len(d1.table) == 5 and d1.trimmed == 10  # reads 2,5 share a centre


def test_quality_trim_known_answers():
    # BWA rule: cutoff 10, qualities (3' end) ... 40 40 5 5 -> trim the two low bases
    q = "".join(chr(33 + v) for v in [40, 40, 40, 5, 5])
    assert po.quality_trim_index(q, 0, 10) == (0, 3)
    # a low base followed by good ones at the very end is kept when the running sum goes negative
    q = "".join(chr(33 + v) for v in [40, 5, 40])
    assert po.quality_trim_index(q, 0, 10) == (0, 3)
    # everything bad -> (0, 0)
    assert po.quality_trim_index("".join(chr(33 + 2) for _ in range(6)), 0, 10) == (0, 0)
    # 5' side
    q = "".join(chr(33 + v) for v in [2, 2, 30, 30])
    assert po.quality_trim_index(q, 10, 10) == (2, 4)
    # NextSeq: high-quality G tail is trimmed
    assert po.nextseq_trim_index("ACGTGGGG", "IIIIIIII", 20) == 4
    assert po.nextseq_trim_index("ACGTGGGA", "IIIIIIII", 20) == 8


def test_adapter_edge_cases():
    ad = po.Adapter("back", ILL_AD)
    ins = "ACGTTGCATGCAAGTCCGTA"
    # 3-nt partial adapter at the very end is accepted (cost 0), 2-nt is not (Appendix A)
    assert po.match_to(ad, ins + ILL_AD[:3]) == (0, 3, len(ins), len(ins) + 3, 3, 0)
    assert po.match_to(ad, ins + ILL_AD[:2]) is None
    # two full occurrences: leftmost wins
    r = ins + ILL_AD + "AC" + ILL_AD
    assert po.match_to(ad, r)[2] == len(ins)
    # adapter at position 0 -> empty read
    assert po.match_to(ad, ILL_AD + "ACGT")[2] == 0
    # N inside the adapter region of the read counts as an error
    r = ins + ILL_AD[:10] + "N" + ILL_AD[11:]
    mt = po.match_to(ad, r)
    assert mt[2] == len(ins) and mt[4:] == (28, 1)
    # empty read
    assert po.match_to(ad, "") is None
    # full DP and find fast path agree
    assert po.locate(ad, ins + ILL_AD + "TTT") == po.match_to(ad, ins + ILL_AD + "TTT")
    # max error table for rate 0.12
    for L, e in [(8, 0), (9, 1), (16, 1), (17, 2), (24, 2), (25, 3), (33, 3), (34, 4)]:
        assert int(L * 0.12) == e
    # front adapter
    g = po.Adapter("front", "GTTCAGAGTTCTACAGTCCGACGATC")
    mt = po.match_to(g, "GTTCAGAGTTCTACAGTCCGACGATC" + ins)
    assert mt[3] == 26
    mt = po.match_to(g, "CCGACGATC" + ins)  # partial 5' adapter (suffix of the adapter) at read start
    assert mt[:4] == (17, 26, 0, 9)
    # case: matching is on upper(), trimming on the original
    assert po.match_to(ad, (ins + ILL_AD).lower())[2] == len(ins)


def test_qiagen_quirks():
    p = po.TrimParams(adapters=[po.Adapter("back", QIA_AD)], umi=(0, 12), qiagenumi=True, minimum_length=0)
    # adapter at position 0: trimmed == "" -> "".split("") ValueError -> umi "" (digest.py:345-346)
    r = QIA_AD + "ACGTACGTACGT" + "AGATCGG"
    assert [k for k, _ in po.digest_read(r, "I" * len(r), p)] == [""]
    # trimmed sequence occurring again later: split()[1] stops at the second occurrence
    t = "ACGTAC"
    r = t + QIA_AD + "GG" + t + "TTTTTTTTTT"
    out = po.digest_read(r, "I" * len(r), p)
    assert out[0][0] == t + r.split(t)[1][: len(QIA_AD) + 12][-12:]


def test_umi_parser_quirk_b0():
    assert po.umi_parser("ACGTACGT", 2, 0) == ("GTACGT", "AC" + "ACGTACGT")  # digest.py:313
    assert po.umi_parser("ACGTACGT", 2, 2) == ("GTAC", "ACGT")
    assert po.umi_parser("ACG", 2, 2) == ("", "ACCG")


def test_read_chunks_and_parse():
    recs = fq(["ACGT" * 5] * 10)
    one = len(recs) // 10
    ch = po.read_chunks(recs, buffer_size=one * 3 + 5)
    assert ch[0] == (0, one * 3) and ch[-1][1] == len(recs)
    assert sum(len(po.parse_fastq(recs[s:e])) for s, e in ch) == 10
    assert po.read_chunks(b"") == []
    # no trailing newline: the tail record is still yielded and parsed
    ch = po.read_chunks(recs[:-1], buffer_size=one * 4)
    assert ch[-1][1] == len(recs) - 1
    assert sum(len(po.parse_fastq(recs[:-1][s:e])) for s, e in ch) == 10
    with pytest.raises(po.FastqFormatError):
        po.parse_fastq(b"@a\nACGT\n+\nIII\n")
    with pytest.raises(po.FastqFormatError):
        po.parse_fastq(b"a\nACGT\n+\nIIII\n")
    assert po.parse_fastq(b"@a\r\nACGT\r\n+\r\nIIII\r\n") == [("a", "ACGT", "IIII")]


def test_round_policies_and_sam_vector():
    # summary.py:1194: AAAACATCAGATTGTGAGTC aligned with one mismatch (MD:Z:17A2) at POS 18 in round 2
    ref = "G" * 17 + "AAAACATCAGATTGTGAATC" + "CCA"
    lib = po.Library(["trnaMT_HisGTG_MT_+_12138_12206"], [ref])
    h = po.hits("AAAACATCAGATTGTGAGTC", lib, po.ROUND_POLICIES[2])
    assert h == [(1, 0, 17, 1)]  # 0-based offset 17 == SAM POS 18, NM:i:1
    assert po.hits("AAAACATCAGATTGTGAGTC", lib, po.ROUND_POLICIES[3]) == []
    # round 3 query rewrite: trailing T-run removed (manifoldAlign.py:122)
    assert po.round_query("ACGTACGTTTTT", 3) == "ACGTACG"
    assert po.round_query("ACGTACGTT", 3) is None
    # round 8: -5 1 -3 2
    assert po.round_query("TGAGGTAGTAGGTTGTATAGTT", 8) == "GAGGTAGTAGGTTGTATAG"
    # -n 1: <=1 mismatch in the first 28, <=2 overall
    ref = "ACGTTGCAAGGCTTAACCGGTTAACGTGCATGCAAGTC"
    lib = po.Library(["r"], [ref])
    q = list(ref[:36])
    q[30] = "A" if q[30] != "A" else "C"
    q[33] = "A" if q[33] != "A" else "C"
    assert po.hits("".join(q), lib, po.ROUND_POLICIES[1]) == [(2, 0, 0, 0)]
    q[3] = "A" if q[3] != "A" else "C"
    assert po.hits("".join(q), lib, po.ROUND_POLICIES[1]) == []
    # reference N may not be overlapped; read N is a mismatch
    lib = po.Library(["r"], ["ACGTNACGTACGTACGTACG"])
    assert po.hits("ACGTACGTACGTACG", lib, po.ROUND_POLICIES[2]) == [(0, 0, 5, 0)]
    assert po.hits("ACGTACNTACGTACG", lib, po.ROUND_POLICIES[2]) == [(1, 0, 5, 1)]
