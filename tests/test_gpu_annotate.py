"""GPU parity of the annotation rounds (mirge_annotate_round through the C ABI) against the oracle's
exhaustive-scan definition of bowtie's valid hit set + canonical pick, and of the two drop-in
entry points (baking, bwtAlign) end to end."""
import argparse
import os

import numpy as np
import pytest

from mirge_b200 import abi
from oracle import coracle, pyoracle as po
from tests.util import ILL, make_args, random_fastq  # noqa: F401

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    from mirge_b200 import device as D

    return D.Device(0)


def make_libs(rng, scale=1.0):
    B = np.array(list("ACGT"))

    def rnd(n, lo, hi):
        return ["".join(rng.choice(B, rng.integers(lo, hi + 1))) for _ in range(int(n * scale))]

    libs = {
        "mirna": rnd(300, 18, 25),
        "hairpin": rnd(200, 60, 120),
        "mature_trna": [s + "CCA" for s in rnd(60, 70, 90)],
        "pre_trna": rnd(80, 90, 150),
        "snorna": rnd(100, 60, 300),
        "rrna": rnd(8, 120, 2000),
        "ncrna_others": rnd(300, 100, 600),
        "mrna": rnd(150, 500, 2000),
        "spike-in": rnd(20, 22, 22),
    }
    # hairpins embed a miRNA; a few reference bases are ambiguous
    for i, h in enumerate(libs["hairpin"]):
        m = libs["mirna"][i % len(libs["mirna"])]
        o = int(rng.integers(0, len(h) - len(m)))
        libs["hairpin"][i] = h[:o] + m + h[o + len(m):]
    for k in ("ncrna_others", "mrna"):
        for i in range(0, len(libs[k]), 7):
            s = libs[k][i]
            p = int(rng.integers(len(s)))
            libs[k][i] = s[:p] + "N" + s[p + 1:]
    return {k: (["%s_%d" % (k, i) for i in range(len(v))], v) for k, v in libs.items()}


def make_queries(rng, libs, n):
    B = np.array(list("ACGT"))
    keys = list(libs)
    out = set()
    while len(out) < n:
        k = keys[int(rng.integers(len(keys)))]
        ref = libs[k][1][int(rng.integers(len(libs[k][1])))]
        L = int(rng.integers(16, 45))
        a = int(rng.integers(0, max(1, len(ref) - L + 1)))
        s = list(ref[a:a + L].replace("N", "A"))
        r = rng.random()
        nm = 0 if r < 0.4 else (1 if r < 0.7 else (2 if r < 0.9 else 3))
        for _ in range(nm):
            s[int(rng.integers(len(s)))] = str(rng.choice(B))
        if rng.random() < 0.1:
            s += list("T" * int(rng.integers(3, 7)))
        if rng.random() < 0.15:  # isomiR-like: extra bases at both ends
            s = [str(rng.choice(B))] + s + list(rng.choice(B, 2))
        if rng.random() < 0.03:
            s[int(rng.integers(len(s)))] = "N"
        if rng.random() < 0.02:
            j = int(rng.integers(len(s)))
            s[j] = s[j].lower()
        if rng.random() < 0.05:
            s = list(rng.choice(B, int(rng.integers(13, 40))))
        out.add("".join(s))
    return sorted(out)


def oracle_annotate(seqs, libs, spike):
    keys = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    ar = np.full(len(seqs), 0xFF, dtype=np.uint8)
    hit = np.full(len(seqs), abi.NO_HIT, dtype=np.uint64)
    for rnd in range(10 if spike else 9):
        pol = po.ROUND_POLICIES[rnd]
        sel = 0 if rnd == 0 else (1 if rnd == 1 else 2)
        cpol = abi.RoundPolicy(rnd, sel, pol.seed_len, pol.seed_mm, pol.total_mm, pol.trim5, pol.trim3, int(pol.strip_polyT))
        names, refs = libs[po.ROUND_LIBS[rnd]]
        rtext = np.frombuffer("".join(refs).upper().encode(), dtype=np.uint8)
        roff = np.zeros(len(refs) + 1, dtype=np.uint32)
        roff[1:] = np.cumsum([len(r) for r in refs])
        coracle.annotate_round(keys, off, rtext, roff, cpol, ar, hit, nthreads=8)
    return ar, hit


def test_rounds_match_oracle(dev):
    from mirge_b200 import libraries as LB
    from mirge_b200 import manifoldAlign as MA

    rng = np.random.default_rng(5)
    libs = make_libs(rng)
    seqs = make_queries(rng, libs, 6000)
    ls = LB.LibrarySet.from_fasta_dict(dev, {k: (n, [s.encode() for s in v]) for k, (n, v) in libs.items()})
    ks = MA.KeySet.from_strings(dev, seqs)
    for spike in (False, True):
        a_g, h_g = MA.annotate_keys(dev, ls, ks, spike)
        a_g = a_g.cpu().numpy()
        h_g = h_g.cpu().numpy().view(np.uint64)
        a_o, h_o = oracle_annotate(seqs, libs, spike)
        bad = np.nonzero((a_g != a_o) | (h_g != h_o))[0]
        assert bad.size == 0, "first mismatch: %s gpu=(%d,%x) oracle=(%d,%x)" % (
            seqs[bad[0]], a_g[bad[0]], h_g[bad[0]], a_o[bad[0]], h_o[bad[0]])
        # every round annotated something, and some sequences stay unannotated
        assert set(np.unique(a_o)) >= set(range(9)), np.unique(a_o)
        assert (a_o == 0xFF).sum() > 0
        # one launch per round (mirge_annotate_round) gives the same as the fused launch
        # (the default above is the split form: filter pre-pass + one list-driven search per round)
        for kw in ({"fused": False}, {"fused": True, "split": False}):
            a_u, h_u = MA.annotate_keys(dev, ls, ks, spike, **kw)
            assert np.array_equal(a_u.cpu().numpy(), a_g) and np.array_equal(h_u.cpu().numpy().view(np.uint64), h_g), kw


def test_python_oracle_spot_check(dev):
    """A few hundred sequences against the pure-Python restatement as well (independent of the C oracle)."""
    from mirge_b200 import libraries as LB
    from mirge_b200 import manifoldAlign as MA

    rng = np.random.default_rng(9)
    libs = make_libs(rng, scale=0.15)
    seqs = make_queries(rng, libs, 250)
    ls = LB.LibrarySet.from_fasta_dict(dev, {k: (n, [s.encode() for s in v]) for k, (n, v) in libs.items()})
    a_g, h_g = MA.annotate_keys(dev, ls, MA.KeySet.from_strings(dev, seqs), True)
    a_g = a_g.cpu().numpy()
    h_g = h_g.cpu().numpy().view(np.uint64)
    pl = {k: po.Library(n, [s.upper() for s in v]) for k, (n, v) in libs.items()}
    exp = po.annotate(seqs, pl, spike_in=True)
    for i, s in enumerate(seqs):
        if s in exp:
            rnd, name, off, mm = exp[s]
            h = int(h_g[i])
            assert a_g[i] == rnd and pl[po.ROUND_LIBS[rnd]].names[(h >> 28) & 0xFFFFFFF] == name
            assert (h & 0xFFFFFFF, h >> 56) == (off, mm)
        else:
            assert a_g[i] == 0xFF


def write_lib_dir(tmp, libs, organism="human", db="miRBase"):
    from mirge_b200.libraries import INDEX_SUFFIX, ROUND_LIBS

    d = os.path.join(tmp, organism, "index.Libs")
    os.makedirs(d, exist_ok=True)
    for rnd, key in enumerate(ROUND_LIBS):
        name = organism + INDEX_SUFFIX[rnd] + (db if rnd in (0, 1, 8) else "")
        names, seqs = libs[key]
        with open(os.path.join(d, name + ".fa"), "w") as f:
            for n, s in zip(names, seqs):
                f.write(">%s some description\n" % n)
                for i in range(0, len(s), 60):
                    f.write(s[i:i + 60] + "\n")
    return tmp


def test_baking_and_bwtalign_end_to_end(dev, tmp_path):
    """Two samples through the reference-facing entry points; compared with the oracle's restatement of
    baking (per-sample dicts + outer join) and bwtAlign (round loop)."""
    from mirge_b200 import digest as DG
    from mirge_b200 import manifoldAlign as MA

    rng = np.random.default_rng(21)
    libs = make_libs(rng, scale=0.5)
    write_lib_dir(str(tmp_path / "lib"), libs)
    mir = libs["mirna"][1]
    files, names, exp_tabs = [], [], []
    for si in range(2):
        # reads whose inserts are library miRNAs so that annotation has something to find
        data = bytearray()
        r2 = np.random.default_rng(100 + si)
        for i in range(3000):
            ins = mir[int(r2.zipf(1.4)) % len(mir)]
            if r2.random() < 0.2:
                ins = ins[1:] + "A"
            s = (ins + ILL + "ACGTACGTAC")[:50]
            q = "".join(chr(33 + int(v)) for v in np.clip(r2.integers(20, 41, len(s)), 0, 41))
            data += ("@s%d.%d\n%s\n+\n%s\n" % (si, i, s, q)).encode()
        p = tmp_path / ("sample%d.fastq" % si)
        p.write_bytes(bytes(data))
        files.append(str(p))
        names.append("sample%d" % si)
    args = make_args(libraries_path=str(tmp_path / "lib"), spikeIn=True)
    df, src, trc, tru = DG.baking(args, files, names, str(tmp_path), device=dev, batch_bytes=100000)
    # oracle side
    from mirge_b200 import params as P

    cp = P.build_trim_params(P.TrimConfig.from_args(args))
    union = {}
    for si, f in enumerate(files):
        raw = np.fromfile(f, dtype=np.uint8)
        n, tab = coracle.digest_collapse(raw, cp, nthreads=4)
        d = tab.to_dict()
        assert src[names[si]] == n == 3000
        assert trc[names[si]] == sum(d.values()) and tru[names[si]] == len(d)
        for k, c in d.items():
            union.setdefault(k, [0, 0])[si] = c
    assert list(df.index) == sorted(union)
    assert list(df.columns) == ["annotFlag"] + DG.INITIAL_FLAGS + names
    for si in range(2):
        assert df[names[si]].tolist() == [union[k][si] for k in df.index]
    assert str(df["annotFlag"].dtype) == "int64" and str(df[names[0]].dtype) == "int64"
    out = MA.bwtAlign(args, df, str(tmp_path), "miRBase", device=dev)
    seqs = list(out.index)
    a_o, h_o = oracle_annotate(seqs, libs, True)
    for rnd in range(10):
        col = out[po.ROUND_COLUMNS[rnd]].tolist()
        nm = libs[po.ROUND_LIBS[rnd]][0]
        for i, s in enumerate(seqs):
            exp = nm[(int(h_o[i]) >> 28) & 0xFFFFFFF] if a_o[i] == rnd else ""
            assert col[i] == exp, (rnd, s, col[i], exp)
    assert out["annotFlag"].tolist() == [int(x != 0xFF) for x in a_o]
    assert (out["annotFlag"] == 1).sum() > 10
    log = (tmp_path / "run.log").read_text()
    assert "Data pre-processing completed" in log and "Alignment completed" in log
    # without -spk the spike-in column is dropped (manifoldAlign.py:137-138)
    args2 = make_args(libraries_path=str(tmp_path / "lib"), spikeIn=False)
    df2, *_ = DG.baking(args2, files[:1], names[:1], str(tmp_path), device=dev)
    out2 = MA.bwtAlign(args2, df2, str(tmp_path), "miRBase", device=dev)
    assert "spike-in" not in out2.columns


def test_streamed_annotation_matches_whole_table(dev):
    """StreamedAnnotator (annotate + D2H the keys each piece created) == annotating the finished table."""
    from mirge_b200 import device as D
    from mirge_b200 import digest as DG
    from mirge_b200 import libraries as LB
    from mirge_b200 import manifoldAlign as MA
    from mirge_b200 import params as P
    from mirge_b200 import synth

    libs = synth.make_libraries(scale=0.02, mrna_count=30)
    lset = LB.LibrarySet.from_fasta_dict(dev, libs.fasta_dict())
    eng = D.DigestEngine(dev, synth.trim_config_for(2, "head"))
    fq = synth.ReadGenerator(libs, synth.CONFIGS[2], dev.tdev).fastq(20000)
    host = fq.cpu().pin_memory()
    table = D.CollapseTable(dev, min_keys=1 << 10)
    sa = MA.StreamedAnnotator(dev, lset, True)
    n = DG.HostStreamer(eng, 300_000, max_record=4096).run(host, table, on_piece=sa)
    n_pairs = sa.finish(table)
    assert n == 20000 and sa.n_done == table.n_keys == n_pairs
    annot, hit = MA.annotate_keys(dev, lset, MA.KeySet.from_table(table), True)
    nk, nw = table.n_keys, table.arena_used
    assert np.array_equal(sa.host["annot"][:nk].numpy(), annot.cpu().numpy())
    assert np.array_equal(sa.host["hit"][:nk].numpy(), hit.cpu().numpy())
    assert (annot.cpu().numpy() != 0xFF).sum() > 100
    assert np.array_equal(sa.host["key_ref"][:nk].numpy(), table.key_ref[:nk].cpu().numpy())
    # arena words referenced by keys arrived intact (dead space of duplicate emissions is copied too)
    assert np.array_equal(sa.host["arena"][:nw].numpy(), table.arena[:nw].cpu().numpy())
    # counts: same multiset as a resident pass; the device-only variant (digest_device hook) gives the same arrays
    t2 = D.CollapseTable(dev, min_keys=1 << 10)
    sb = MA.StreamedAnnotator(dev, lset, True, to_host=False)
    eng.digest_device(fq, t2, 250_000, on_piece=sb)
    ids2, cnt2 = t2.drain()
    assert sorted(sa.host["cnt"][:n_pairs].tolist()) == sorted(cnt2.cpu().tolist())
    a2, h2 = sb.results()
    ann2, hit2 = MA.annotate_keys(dev, lset, MA.KeySet.from_table(t2), True)
    assert torch.equal(a2, ann2) and torch.equal(h2, hit2) and not sb.host


def test_annotation_with_the_full_size_libraries_matches_oracle(dev):
    """The nine rounds on > 1 M unique sequences of the C2 workload against the libraries at the size bench.py uses --
    all of them at full scale, the mRNA library with its 100 000 entries (150 M bases: index, bucket table and presence
    filters at the sizes the timed kernels see) -- hit by hit against the oracle's indexed CPU search."""
    from mirge_b200 import device as D
    from mirge_b200 import libraries as LB
    from mirge_b200 import manifoldAlign as MA
    from mirge_b200 import synth

    libs = synth.make_libraries(mrna_count=100_000)
    lset = LB.LibrarySet.from_fasta_dict(dev, libs.fasta_dict())
    eng = D.DigestEngine(dev, synth.trim_config_for(2, "head"))
    fq = synth.ReadGenerator(libs, synth.CONFIGS[2], dev.tdev).fastq(1_300_000)
    table = D.CollapseTable(dev, min_keys=1 << 21)
    assert eng.digest_device(fq, table, 128 << 20) == 1_300_000
    nk = table.n_keys
    assert nk > 1_000_000
    annot, hit = MA.annotate_keys(dev, lset, MA.KeySet.from_table(table), False)
    annot, hit = annot.cpu().numpy(), hit.cpu().numpy().view(np.uint64)
    kk = table.export_keys()
    blob = np.frombuffer(b"".join(kk.tolist()), dtype=np.uint8)
    off = np.zeros(len(kk) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(np.char.str_len(kk))
    a_o = np.full(nk, 0xFF, dtype=np.uint8)
    h_o = np.full(nk, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    pols = LB.round_policies()
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
    for rnd in range(9):
        L = libs.libs[LB.ROUND_LIBS[rnd]]
        index = coracle.Index(lut[L.codes], L.off.astype(np.uint32))
        coracle.annotate_round_indexed(blob, off, index, pols[rnd], a_o, h_o, threads)
        del index
    bad = np.flatnonzero((annot != a_o) | (hit != h_o))
    assert bad.size == 0, "first differing sequence %r: gpu round %d hit %x, oracle round %d hit %x (%d differ)" % (
        kk[bad[0]], annot[bad[0]], hit[bad[0]], a_o[bad[0]], h_o[bad[0]], bad.size)
    per_round = np.bincount(a_o[a_o != 0xFF], minlength=9)
    assert (per_round > 1000).all(), per_round  # every round, the mRNA round included, annotated thousands of sequences


def test_libraries_load_from_bowtie_index_files(dev, tmp_path, monkeypatch):
    """A library directory that ships only .ebwt files (stock miRge3_Lib) gives the same annotation as FASTA -- once the
    built-in decoder is asked for: without MIRGE_B200_TRUST_EBWT=1 (and without a bowtie-inspect) the loader refuses."""
    monkeypatch.setenv("PATH", str(tmp_path))  # no bowtie-inspect
    monkeypatch.delenv("MIRGE_B200_TRUST_EBWT", raising=False)
    from mirge_b200 import libraries as LB
    from mirge_b200 import manifoldAlign as MA
    from mirge_b200.libraries import INDEX_SUFFIX, ROUND_LIBS
    from tests.util import write_ebwt

    rng = np.random.default_rng(33)
    libs = make_libs(rng, scale=0.3)
    write_lib_dir(str(tmp_path / "fa"), libs)
    d = tmp_path / "eb" / "human" / "index.Libs"
    d.mkdir(parents=True)
    for rnd, key in enumerate(ROUND_LIBS):
        names, seqs = libs[key]
        write_ebwt(str(d / ("human" + INDEX_SUFFIX[rnd] + ("miRBase" if rnd in (0, 1, 8) else ""))),
                   ["%s some description" % n for n in names], seqs)
    a = LB.LibrarySet.from_mirge_lib(dev, str(tmp_path / "fa"), "human", "miRBase", True)
    from mirge_b200.device import MirgeError

    with pytest.raises(MirgeError, match="MIRGE_B200_TRUST_EBWT"):
        LB.LibrarySet.from_mirge_lib(dev, str(tmp_path / "eb"), "human", "miRBase", True)
    monkeypatch.setenv("MIRGE_B200_TRUST_EBWT", "1")
    with pytest.warns(RuntimeWarning, match="not yet bowtie-validated"):
        b = LB.LibrarySet.from_mirge_lib(dev, str(tmp_path / "eb"), "human", "miRBase", True)
    seqs = make_queries(rng, libs, 1500)
    ks = MA.KeySet.from_strings(dev, seqs)
    ra, ha = MA.annotate_keys(dev, a, ks, True)
    rb, hb = MA.annotate_keys(dev, b, ks, True)
    assert torch.equal(ra, rb) and torch.equal(ha, hb) and int((ra != 0xFF).sum()) > 200
    for key in set(ROUND_LIBS):
        assert a[key].names == b[key].names
