"""The C oracle must agree with the readable Python restatement (pyoracle) window by window."""
import zlib

import numpy as np
import pytest

from mirge_b200 import abi
from mirge_b200 import params as P
from oracle import coracle, pyoracle as po
from tests.util import CONFIG_DATA, CONFIGS, py_params, random_fastq


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_c_matches_python(name):
    cfg = CONFIGS[name]
    data = random_fastq(400, seed=zlib.crc32(name.encode()) % 1000, n_rate=0.02, lower_rate=0.01,
                        **CONFIG_DATA.get(name, {}))
    fq = np.frombuffer(data, dtype=np.uint8)
    cp = P.build_trim_params(cfg)
    pp = py_params(cfg)
    E = P.trim_slots(cp)
    n, win, kept = coracle.trim(fq, cp)
    recs = po.parse_fastq(data)
    assert n == len(recs)
    pydict = {}
    for r, (_nm, seq, qual) in enumerate(recs):
        keys_py = [k for k, _ in po.digest_read(seq, qual, pp)]
        keys_c = [seq[win[r, s, 0] : win[r, s, 1]] + seq[win[r, s, 2] : win[r, s, 3]] for s in range(E) if kept[r, s]]
        assert keys_c == keys_py, (r, seq)
        for k in keys_py:
            pydict[k] = pydict.get(k, 0) + 1
    for nt in (1, 3):
        n2, tab = coracle.digest_collapse(fq, cp, nthreads=nt)
        assert n2 == n and tab.to_dict() == pydict and tab.total == sum(pydict.values())
    umi = cfg.umi()
    if umi is not None:
        for dedup in (False, True):
            d = po.digest_sample(data, pp, umi_dedup=dedup)
            t2 = tab.umi_collapse(umi[0], umi[1], cfg.minimum_length, dedup)
            assert t2.to_dict() == d.table and t2.total == d.trimmed


def test_u_in_reads_is_t_only_through_the_wildcard_tables():
    """cutadapt compares characters when no wildcard is active (a U in the read is not a T) and translated sets when the
    adapter or the read may hold wildcards (U = T): both oracles, on reads that carry the adapter with U for T."""
    rng = np.random.default_rng(5)
    B = np.array(list("ACGT"))
    for name in ("default", "wild"):
        cfg = CONFIGS[name]
        ad = cfg.adapters[0][1].replace("N", "A").replace("R", "G")
        recs, n_u = [], 0
        for i in range(300):
            s = "".join(rng.choice(B, int(rng.integers(16, 30)))) + ad
            if i % 2:
                s = s.replace("T", "U")
                n_u += 1
            recs.append("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)))
        data = "".join(recs).encode()
        cp, pp = P.build_trim_params(cfg), py_params(cfg)
        n, win, kept = coracle.trim(np.frombuffer(data, dtype=np.uint8), cp)
        trimmed_u = 0
        for r, (_nm, seq, qual) in enumerate(po.parse_fastq(data)):
            w_py = po.digest_read(seq, qual, pp)[-1][1]
            assert tuple(int(x) for x in win[r, -1]) == tuple(w_py), (name, seq)
            trimmed_u += "U" in seq and w_py[1] < len(seq)
        # plain adapter: the U reads keep their adapter (apart from chance partial matches); wildcard adapter: all are trimmed
        assert (trimmed_u == n_u) if name == "wild" else (trimmed_u < n_u // 4), (name, trimmed_u, n_u)


def test_crlf_and_no_final_newline():
    cfg = CONFIGS["default"]
    cp = P.build_trim_params(cfg)
    a = random_fastq(50, seed=5)
    b = random_fastq(50, seed=5, crlf=True)
    c = random_fastq(50, seed=5, final_newline=False)
    ref = coracle.digest_collapse(np.frombuffer(a, dtype=np.uint8), cp)[1].to_dict()
    for d in (b, c):
        assert coracle.digest_collapse(np.frombuffer(d, dtype=np.uint8), cp)[1].to_dict() == ref
    assert po.digest_sample(b, py_params(cfg)).table == ref


def test_format_errors():
    cp = P.build_trim_params(CONFIGS["default"])
    for bad in (b"@a\nACGT\n+\nIII\n", b"@a\nACGT\n+\n", b"a\nACGT\n+\nIIII\n", b"@a\nACGT\n-\nIIII\n"):
        with pytest.raises(ValueError):
            coracle.trim(np.frombuffer(bad, dtype=np.uint8), cp)
        with pytest.raises(po.FastqFormatError):
            po.parse_fastq(bad)
    n, win, kept = coracle.trim(np.frombuffer(b"", dtype=np.uint8), cp)
    assert n == 0


def test_annotate_round_c_vs_python():
    rng = np.random.default_rng(3)
    B = np.array(list("ACGT"))
    refs = ["".join(rng.choice(B, rng.integers(18, 80))) for _ in range(60)]
    refs[5] = refs[5][:10] + "N" + refs[5][11:]
    lib = po.Library(["ref%d" % i for i in range(len(refs))], refs)
    seqs = []
    for i in range(300):
        r = refs[int(rng.integers(len(refs)))]
        a = int(rng.integers(0, max(1, len(r) - 16)))
        s = list(r[a : a + int(rng.integers(13, 40))])
        for _ in range(int(rng.integers(0, 4))):
            if s:
                s[int(rng.integers(len(s)))] = str(rng.choice(B))
        if rng.random() < 0.2:
            s += list("TTTT")
        if rng.random() < 0.05 and s:
            s[int(rng.integers(len(s)))] = "N"
        seqs.append("".join(s))
    seqs = sorted(set(s for s in seqs if s))
    keys = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    rtext = np.frombuffer("".join(refs).encode(), dtype=np.uint8)
    roff = np.zeros(len(refs) + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(r) for r in refs])
    nhit = 0
    for rnd, pol in enumerate(po.ROUND_POLICIES):
        sel = 0 if rnd == 0 else (1 if rnd == 1 else 2)
        cpol = abi.RoundPolicy(rnd, sel, pol.seed_len, pol.seed_mm, pol.total_mm, pol.trim5, pol.trim3,
                               int(pol.strip_polyT))
        ar = np.full(len(seqs), 0xFF, dtype=np.uint8)
        hit = np.full(len(seqs), abi.NO_HIT, dtype=np.uint64)
        coracle.annotate_round(keys, off, rtext, roff, cpol, ar, hit, nthreads=2)
        for i, s in enumerate(seqs):
            if (sel == 0 and not len(s) < 26) or (sel == 1 and not len(s) > 25):
                exp = None
            else:
                q = po.round_query(s, rnd)
                exp = po.canonical_pick(po.hits(q, lib, pol)) if q is not None else None
            if exp is None:
                assert ar[i] == 0xFF, (rnd, s)
            else:
                nhit += 1
                h = int(hit[i])
                assert ar[i] == rnd and (h >> 56, (h >> 28) & 0xFFFFFFF, h & 0xFFFFFFF) == exp, (rnd, s)
    assert nhit > 200


def test_indexed_search_equals_exhaustive_scan():
    """The indexed CPU search (used as the CPU baseline on large libraries) must return exactly what the
    exhaustive scan -- the definition -- returns, round by round."""
    from mirge_b200 import synth

    libs = synth.make_libraries(scale=0.03, mrna_count=40)
    gen = synth.ReadGenerator(libs, synth.CONFIGS[1], "cpu")
    fq = gen.fastq(4000).numpy()
    cp = P.build_trim_params(synth.trim_config_for(1))
    _, tab = coracle.digest_collapse(fq, cp, nthreads=4)
    keys, off, cnt = tab.export()
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    from mirge_b200.libraries import ROUND_LIBS, round_policies

    pols = round_policies()
    a1 = np.full(len(cnt), 0xFF, dtype=np.uint8)
    h1 = np.full(len(cnt), abi.NO_HIT, dtype=np.uint64)
    a2, h2 = a1.copy(), h1.copy()
    for rnd in range(10):
        L = libs.libs[ROUND_LIBS[rnd]]
        text, roff = lut[L.codes], L.off.astype(np.uint32)
        coracle.annotate_round(keys, off, text, roff, pols[rnd], a1, h1, nthreads=8)
        coracle.annotate_round_indexed(keys, off, coracle.Index(text, roff), pols[rnd], a2, h2, nthreads=8)
        assert np.array_equal(a1, a2) and np.array_equal(h1, h2), rnd
    assert (a1 != 0xFF).sum() > 100 and len(set(a1.tolist())) >= 6


def test_synthetic_generator_is_seeded_and_well_formed():
    from mirge_b200 import synth

    libs = synth.make_libraries(scale=0.02, mrna_count=20)
    for cid in (1, 2, 3):
        a = synth.ReadGenerator(libs, synth.CONFIGS[cid], "cpu").fastq(3000).numpy().tobytes()
        b = synth.ReadGenerator(libs, synth.CONFIGS[cid], "cpu").fastq(3000).numpy().tobytes()
        assert a == b
        recs = po.parse_fastq(a)
        assert len(recs) == 3000 and all(len(s) == synth.CONFIGS[cid].L == len(q) for _, s, q in recs)
        assert recs[12][0].startswith("SYN%d.12 12 length=" % cid)
    # most C1 reads carry the (possibly partial) illumina adapter
    ad = po.Adapter("back", synth.ILLUMINA)
    recs = po.parse_fastq(synth.ReadGenerator(libs, synth.CONFIGS[1], "cpu").fastq(300).numpy().tobytes())
    assert sum(po.match_to(ad, s) is not None for _, s, _ in recs) > 250
