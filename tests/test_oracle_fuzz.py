"""Seeded fuzz of the two oracle implementations against each other: random trim configurations (adapter text with
IUPAC wildcards, 3' / 5' / both, error rate, minimum overlap, indels on / off, removal rounds, NextSeq and two-sided
quality cut-offs, -NX, cuts, minimum length, UMI flanks, both counting modes) times reads built to provoke the search
(adapter copies with substitutions and indels, partial adapters at the 3' end, repeats, poly-G tails, N, lower case).
The readable Python restatement (pinned on the reference-written golden files) and the C port (what the GPU path is
held to at scale) must give the same emitted keys read by read."""
import numpy as np
import pytest

from mirge_b200 import params as P
from oracle import coracle, pyoracle as po
from tests.util import py_params

B = np.array(list("ACGT"))
IUPAC = "RYSWKMBDHVN"


def rnd_seq(rng, n):
    return "".join(rng.choice(B, n))


def mutate(rng, s, p_sub, p_indel):
    out = []
    for c in s:
        r = rng.random()
        if r < p_indel / 2:
            continue
        if r < p_indel:
            out.append(str(rng.choice(B)))
        out.append(str(rng.choice(B)) if rng.random() < p_sub else c)
    return "".join(out)


def random_config(rng):
    n_ad = int(rng.integers(1, 3))
    adapters = []
    for k in range(n_ad):
        m = int(rng.integers(6, 33))
        s = rnd_seq(rng, m)
        if rng.random() < 0.3:  # a few wildcard positions
            s = list(s)
            for _ in range(int(rng.integers(1, 4))):
                s[int(rng.integers(m))] = IUPAC[int(rng.integers(len(IUPAC)))]
            s = "".join(s)
        if rng.random() < 0.15:  # low-complexity adapter: many equally good alignments, tie rules decide
            s = (s[:3] * 12)[:m]
        adapters.append(("back" if (k == 0 and rng.random() < 0.8) or rng.random() < 0.5 else "front", s))
    q = None
    r = rng.random()
    if r < 0.4:
        q = str(int(rng.integers(5, 31)))
    elif r < 0.6:
        q = "%d,%d" % (int(rng.integers(0, 25)), int(rng.integers(0, 31)))
    umi = None
    if rng.random() < 0.2:
        umi = "%d,%d" % (int(rng.integers(0, 7)), int(rng.integers(0, 7)))
    cut = []
    if rng.random() < 0.3:
        cut = [int(rng.integers(1, 5))] if rng.random() < 0.5 else [int(rng.integers(1, 5)), -int(rng.integers(1, 5))]
    backs = [k for k, (w, _) in enumerate(adapters) if w == "back"]
    if backs and rng.random() < 0.2:  # a linked pair -g "A...B": a random 5' half in front of one of the 3' adapters
        adapters[backs[0]] = ("front", rnd_seq(rng, int(rng.integers(5, 14))) + "..." + adapters[backs[0]][1])
    return P.TrimConfig(adapters=adapters, error_rate=float(rng.choice([0.0, 0.05, 0.1, 0.12, 0.2, 0.34])),
                        overlap=int(rng.integers(1, 8)), indels=bool(rng.random() < 0.8), times=int(rng.integers(1, 3)),
                        nextseq_trim=int(rng.integers(10, 31)) if rng.random() < 0.3 else None, quality_cutoff=q,
                        quality_base=33, trim_n=bool(rng.random() < 0.3), cut=cut, minimum_length=int(rng.integers(0, 25)),
                        uniq_mol_ids=umi, count_mode="head" if rng.random() < 0.7 else "release",
                        cutadapt_compat="4" if rng.random() < 0.25 else "2-3")


def random_placement_config(rng):
    """Configurations over the rest of cutadapt's specification language: anchored (^SEQ, SEQ$) and non-internal (XSEQ,
    SEQX) adapters, per-adapter ;parameters, x{n} repeats, linked pairs given with -a / -g with anchored, optional and
    required halves, --match-read-wildcards."""
    def seq(lo=6, hi=25):
        s = rnd_seq(rng, int(rng.integers(lo, hi)))
        if rng.random() < 0.2:
            s = list(s)
            s[int(rng.integers(len(s)))] = IUPAC[int(rng.integers(len(IUPAC)))]
            s = "".join(s)
        if rng.random() < 0.1:
            s = s[:4] + "A{%d}" % int(rng.integers(2, 6)) + s[4:]
        return s

    def prm():
        out = ""
        if rng.random() < 0.25:
            out += ";e=%s" % rng.choice(["0", "0.1", "0.2", "0.3"])
        if rng.random() < 0.2:
            out += ";o=%d" % int(rng.integers(1, 9))
        if rng.random() < 0.15:
            out += ";noindels"
        return out

    def five():
        return str(rng.choice(["%s", "%s", "^%s", "X%s"])) % seq(5, 14)

    def three():
        return str(rng.choice(["%s", "%s", "%s$", "%sX"])) % seq()

    adapters = []
    for k in range(int(rng.integers(1, 3))):
        r = rng.random()
        if r < 0.35:
            adapters.append(("back", three() + prm()))
        elif r < 0.55:
            adapters.append(("front", five() + prm()))
        else:  # linked pair: -g needs both halves, -a only the anchored ones; ;required / ;optional override
            kind = "front" if rng.random() < 0.5 else "back"
            a, b = five() + prm(), three() + prm()
            if rng.random() < 0.25:
                a += str(rng.choice([";optional", ";required"]))
            if rng.random() < 0.25:
                b += str(rng.choice([";optional", ";required"]))
            adapters.append((kind, a + "..." + b))
    q = str(int(rng.integers(5, 31))) if rng.random() < 0.4 else None
    umi = "%d,%d" % (int(rng.integers(0, 5)), int(rng.integers(0, 5))) if rng.random() < 0.15 else None
    return P.TrimConfig(adapters=adapters, error_rate=float(rng.choice([0.0, 0.1, 0.12, 0.2])),
                        overlap=int(rng.integers(1, 8)), indels=bool(rng.random() < 0.8), times=int(rng.integers(1, 3)),
                        nextseq_trim=int(rng.integers(10, 31)) if rng.random() < 0.2 else None, quality_cutoff=q,
                        trim_n=bool(rng.random() < 0.3), cut=[int(rng.integers(1, 4))] if rng.random() < 0.2 else [],
                        minimum_length=int(rng.integers(0, 20)), uniq_mol_ids=umi,
                        match_read_wildcards=bool(rng.random() < 0.35), count_mode="head" if rng.random() < 0.7 else "release",
                        cutadapt_compat="4" if rng.random() < 0.2 else "2-3")


def random_kit_config(rng):
    """More adapters than the bit-parallel kernels take (a kit's list through file:): 5-12 of them, mostly plain 3'
    adapters, some placed or with parameters -- the choice among many matches (most matches, fewest errors, first)."""
    n = int(rng.integers(5, 13))
    adapters = []
    for k in range(n):
        s = rnd_seq(rng, int(rng.integers(8, 30)))
        r = rng.random()
        if r < 0.6:
            adapters.append(("back", s))
        elif r < 0.75:
            adapters.append(("front", s[:12]))
        elif r < 0.85:
            adapters.append(("back", s + str(rng.choice(["$", "X"]))))
        else:
            adapters.append(("back", s + ";e=%s;o=%d" % (rng.choice(["0.05", "0.2"]), int(rng.integers(2, 8)))))
    return P.TrimConfig(adapters=adapters, error_rate=float(rng.choice([0.05, 0.12, 0.2])), overlap=int(rng.integers(2, 7)),
                        indels=bool(rng.random() < 0.8), times=int(rng.integers(1, 4)),
                        quality_cutoff=str(int(rng.integers(5, 25))) if rng.random() < 0.5 else None,
                        minimum_length=int(rng.integers(0, 20)), count_mode="head" if rng.random() < 0.7 else "release")


def random_reads(rng, cfg, n):
    recs = []
    flat = []  # (5' or 3' form, sequence) of every adapter and half
    for kind, spec in cfg.adapters:
        for sp in P.parse_adapter_specs(kind, spec):
            if sp.where == "linked":
                flat += [("front", sp.sequence), ("back", sp.sequence2)]
            else:
                flat.append(("front" if sp.where in ("front", "prefix", "front_not_internal") else "back", sp.sequence))
    for i in range(n):
        L = int(rng.integers(1, 101)) if rng.random() < 0.9 else int(rng.integers(101, 160))
        ins = rnd_seq(rng, int(rng.integers(0, 45)))
        s = ins
        for where, ad in flat:
            plain = "".join(c if c in "ACGT" else str(rng.choice(B)) for c in ad)
            r = rng.random()
            if r < 0.55:
                piece = mutate(rng, plain, 0.06, 0.04 if rng.random() < 0.5 else 0.0)
                s = piece + s if where == "front" and rng.random() < 0.7 else s + piece
            elif r < 0.7:
                s = s + plain[: int(rng.integers(1, len(plain) + 1))]  # partial adapter (ends the read if not padded)
            elif r < 0.8:
                s = s + plain + rnd_seq(rng, 3) + plain  # two occurrences: leftmost-best rule
        if rng.random() < 0.6:
            s = s + rnd_seq(rng, int(rng.integers(0, 40)))
        if rng.random() < 0.15:
            s = s + "G" * int(rng.integers(1, 30))
        s = s[:L] if s else "A"
        s = list(s)
        for j in range(len(s)):
            r = rng.random()
            if r < 0.01:
                s[j] = "N"
            elif r < 0.015:
                s[j] = s[j].lower()
        if rng.random() < 0.1:
            s[0] = "N"
            s[-1] = "N"
        s = "".join(s)
        q = np.clip(rng.integers(20, 42, len(s)) - (np.arange(len(s)) * rng.integers(0, 40) // max(len(s), 1)), 0, 41)
        if rng.random() < 0.1:
            q[:] = rng.integers(0, 42, len(s))
        recs.append("@f%d some text\n%s\n+\n%s\n" % (i, s, "".join(chr(33 + int(v)) for v in q)))
    return "".join(recs).encode()


@pytest.mark.parametrize("seed", range(60))
def test_c_oracle_equals_python_oracle_on_random_configurations(seed):
    rng = np.random.default_rng(9000 + seed)
    cfg = random_config(rng)
    try:
        cp = P.build_trim_params(cfg)
    except P.UnsupportedAdapterSpec:
        pytest.skip("configuration the product rejects")
    pp = py_params(cfg)
    data = random_reads(rng, cfg, 250)
    fq = np.frombuffer(data, dtype=np.uint8)
    E = P.trim_slots(cp)
    n, win, kept = coracle.trim(fq, cp)
    recs = po.parse_fastq(data)
    assert n == len(recs) == 250
    pydict = {}
    for r, (_nm, seq, qual) in enumerate(recs):
        keys_py = [k for k, _ in po.digest_read(seq, qual, pp)]
        keys_c = [seq[win[r, s, 0] : win[r, s, 1]] + seq[win[r, s, 2] : win[r, s, 3]] for s in range(E) if kept[r, s]]
        assert keys_c == keys_py, (seed, cfg, r, seq, qual)
        for k in keys_py:
            pydict[k] = pydict.get(k, 0) + 1
    n2, tab = coracle.digest_collapse(fq, cp, nthreads=2)
    assert n2 == n and tab.to_dict() == pydict
    umi = cfg.umi()
    if umi is not None:
        for dedup in (False, True):
            d = po.digest_sample(data, pp, umi_dedup=dedup)
            t2 = tab.umi_collapse(umi[0], umi[1], cfg.minimum_length, dedup)
            assert t2.to_dict() == d.table and t2.total == d.trimmed


@pytest.mark.parametrize("seed", range(40))
def test_c_oracle_equals_python_oracle_on_random_placements(seed):
    """The same for the rest of the specification language (random_placement_config)."""
    rng = np.random.default_rng(9500 + seed)
    cfg = random_placement_config(rng)
    try:
        cp = P.build_trim_params(cfg)
    except (P.UnsupportedAdapterSpec, RuntimeError) as e:
        pytest.skip("configuration the product rejects: %s" % e)
    pp = py_params(cfg)
    data = random_reads(rng, cfg, 250)
    fq = np.frombuffer(data, dtype=np.uint8)
    E = P.trim_slots(cp)
    n, win, kept = coracle.trim(fq, cp)
    recs = po.parse_fastq(data)
    assert n == len(recs) == 250
    changed = 0
    for r, (_nm, seq, qual) in enumerate(recs):
        keys_py = [k for k, _ in po.digest_read(seq, qual, pp)]
        keys_c = [seq[win[r, s, 0] : win[r, s, 1]] + seq[win[r, s, 2] : win[r, s, 3]] for s in range(E) if kept[r, s]]
        assert keys_c == keys_py, (seed, cfg, r, seq, qual)


@pytest.mark.parametrize("seed", range(10))
def test_c_oracle_equals_python_oracle_on_adapter_kits(seed):
    """... and with 5-12 adapters at once (random_kit_config)."""
    rng = np.random.default_rng(9800 + seed)
    cfg = random_kit_config(rng)
    cp = P.build_trim_params(cfg)
    assert cp.n_adapters >= 5
    pp = py_params(cfg)
    data = random_reads(rng, cfg, 250)
    fq = np.frombuffer(data, dtype=np.uint8)
    E = P.trim_slots(cp)
    n, win, kept = coracle.trim(fq, cp)
    used = set()
    for r, (_nm, seq, qual) in enumerate(po.parse_fastq(data)):
        keys_py = [k for k, _ in po.digest_read(seq, qual, pp)]
        keys_c = [seq[win[r, s, 0] : win[r, s, 1]] + seq[win[r, s, 2] : win[r, s, 3]] for s in range(E) if kept[r, s]]
        assert keys_c == keys_py, (seed, cfg, r, seq, qual)
        which, mt = po.best_match(pp.adapters, seq, pp.compat)
        if mt is not None:
            used.add(id(which))
    assert len(used) >= 3, (seed, len(used))
