"""The product's adapter search (mirge3.0_b200/csrc/adapter_search.cuh: the literal full-column DP ``locate`` and the
bit-parallel ``locate_fast`` with its dominance rules, closed-form cost-1 tracebacks, insertion chains and on-demand
recompute) compiled for the host and held against the Python restatement of cutadapt's ``Aligner.locate``
(oracle/pyoracle.py::locate, pinned on the reference-written golden files) on adversarial inputs: low-complexity
adapters and reads (every tie rule is exercised), adapter copies with substitutions / insertions / deletions, partial
adapters at the 3' end, several occurrences, sub-windows of the read, N and lower case.  The GPU tests compare whole
pipelines on realistic reads; here it is the matcher alone, tens of thousands of searches, on the CPU."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

import mirge_b200
from mirge_b200 import abi
from mirge_b200 import params as P
from oracle import pyoracle as po
from tests.util import host_harness_flags, CONFIG_DATA, CONFIGS, py_params, random_fastq

HERE = os.path.dirname(os.path.abspath(__file__))
B = np.array(list("ACGT"))


@pytest.fixture(scope="module")
def hs(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hs") / "libadapter_search_host.so")
    subprocess.check_call(["g++"] + host_harness_flags() + ["-std=c++17", "-shared", "-fPIC", "-I", os.path.join(mirge_b200.PACKAGE_DIR, "csrc"),
                           "-o", so, os.path.join(HERE, "adapter_search_harness.cpp")])
    lib = C.CDLL(so)
    lib.hs_set_params.argtypes = [C.POINTER(abi.TrimParams), C.POINTER(C.c_int), C.c_char_p, C.c_int]
    lib.hs_locate.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    return lib


def rnd(rng, n):
    return "".join(rng.choice(B, n))


def mutate(rng, s, p_sub, p_ins, p_del):
    out = []
    for c in s:
        r = rng.random()
        if r < p_del:
            continue
        if r < p_del + p_ins:
            out.append(str(rng.choice(B)))
        out.append(str(rng.choice(B)) if rng.random() < p_sub else c)
    return "".join(out)


def make_adapter(rng, kind):
    m = int(rng.integers(4, 33))
    if kind == "random":
        return rnd(rng, m)
    if kind == "homopolymer":
        return str(rng.choice(B)) * m
    if kind == "repeat":
        u = rnd(rng, int(rng.integers(2, 5)))
        return (u * 20)[:m]
    if kind == "two_blocks":  # AAAA...CCCC: insertions and deletions tie everywhere
        a, b = rng.choice(B, 2, replace=False)
        k = int(rng.integers(1, m))
        return str(a) * k + str(b) * (m - k)
    u = rnd(rng, m)  # "prefix_repeat": the adapter's start occurs again inside it
    k = int(rng.integers(2, max(3, m // 2)))
    return (u[:k] + u[: m - k])[:m]


def make_read(rng, ad, kind):
    L = int(rng.integers(0, 110))
    if kind == 0:  # insert + mutated adapter + tail
        s = rnd(rng, int(rng.integers(0, 50))) + mutate(rng, ad, 0.08, 0.04, 0.04) + rnd(rng, int(rng.integers(0, 30)))
    elif kind == 1:  # partial adapter at the 3' end (possibly mutated)
        s = rnd(rng, int(rng.integers(0, 60))) + mutate(rng, ad[: int(rng.integers(1, len(ad) + 1))], 0.05, 0.03, 0.03)
    elif kind == 2:  # several occurrences, the first one damaged
        s = rnd(rng, int(rng.integers(0, 20))) + mutate(rng, ad, 0.15, 0.05, 0.05) + rnd(rng, int(rng.integers(0, 6))) + ad + \
            rnd(rng, int(rng.integers(0, 10))) + mutate(rng, ad, 0.05, 0.0, 0.0)
    elif kind == 3:  # low complexity read made of the adapter's letters
        letters = np.array(sorted(set(ad)))
        s = "".join(rng.choice(letters, int(rng.integers(1, 100))))
    elif kind == 4:  # shifted / doubled copies: runs of insertions and deletions
        k = int(rng.integers(1, max(2, len(ad) // 2)))
        s = rnd(rng, int(rng.integers(0, 30))) + ad[:k] + ad[:k] + ad[k:] + ad[-k:]
    else:
        s = rnd(rng, L)
    s = s[: int(rng.integers(1, 150))] if s else ""
    if s and rng.random() < 0.15:
        s = list(s)
        for _ in range(int(rng.integers(1, 4))):
            j = int(rng.integers(len(s)))
            s[j] = "N" if rng.random() < 0.5 else s[j].lower()
        s = "".join(s)
    return s


def run_case(hs, cfg, reads, stats):
    cp = P.build_trim_params(cfg)
    pp = py_params(cfg)
    fast_ok = C.c_int(0)
    err = C.create_string_buffer(512)
    assert hs.hs_set_params(C.byref(cp), C.byref(fast_ok), err, 512) == 0, err.value
    out = (C.c_int32 * 4)()
    for read in reads:
        raw = read.encode()
        windows = [(0, len(read))]
        if len(read) > 4:
            a, b = sorted(int(x) for x in np.random.default_rng(len(read)).integers(0, len(read) + 1, 2))
            windows.append((a, b))
        for start, stop in windows:
            for ai, ad in enumerate(pp.adapters):
                exp = po.locate(ad, read[start:stop])
                exp4 = None if exp is None else (exp[2], exp[3], exp[4], exp[5])
                modes = [0] + ([1, 2, 3, 4] if fast_ok.value else [])
                for mode in modes:
                    rc = hs.hs_locate(mode, ai, raw, len(raw), start, stop, out)
                    if rc == -1:
                        continue
                    stats[("mode", mode)] = stats.get(("mode", mode), 0) + 1
                    if rc == 2:
                        assert mode in (2, 4)
                        stats["deferred"] = stats.get("deferred", 0) + 1
                        continue
                    got = tuple(out) if rc == 1 else None
                    assert got == exp4, (mode, cfg.adapters[ai], cfg.error_rate, cfg.overlap, cfg.indels, read, (start, stop), got, exp4)
                    if exp4 is not None:
                        stats["matches"] = stats.get("matches", 0) + 1


@pytest.mark.parametrize("seed", range(32))
def test_bit_parallel_and_generic_search_equal_the_oracle(hs, seed):
    rng = np.random.default_rng(4200 + seed)
    stats = {}
    kinds = ["random", "homopolymer", "repeat", "two_blocks", "prefix_repeat"]
    for rep in range(6):
        ad = make_adapter(rng, kinds[(seed + rep) % len(kinds)])
        cfg = P.TrimConfig(adapters=[("back", ad)], error_rate=float(rng.choice([0.0, 0.05, 0.1, 0.12, 0.2, 0.25])),
                           overlap=int(rng.integers(1, 7)), indels=True)
        reads = [make_read(rng, ad, int(rng.integers(0, 6))) for _ in range(60)]
        run_case(hs, cfg, [r for r in reads if r], stats)
    assert stats.get(("mode", 1), 0) > 100 and stats.get("matches", 0) > 100, stats
    if seed == 0:
        print(stats)


@pytest.mark.parametrize("seed", range(6))
def test_generic_search_with_front_adapters_wildcards_and_no_indels(hs, seed):
    rng = np.random.default_rng(4300 + seed)
    stats = {}
    for rep in range(4):
        ad1, ad2 = make_adapter(rng, "random"), make_adapter(rng, "repeat")
        if rng.random() < 0.5:
            ad1 = list(ad1)
            ad1[int(rng.integers(len(ad1)))] = "N"
            ad1[int(rng.integers(len(ad1)))] = "R"
            ad1 = "".join(ad1)
        cfg = P.TrimConfig(adapters=[("front" if rng.random() < 0.5 else "back", ad1), ("front", ad2)],
                           error_rate=float(rng.choice([0.0, 0.1, 0.2])), overlap=int(rng.integers(1, 6)),
                           indels=bool(rng.random() < 0.5))
        plain = "".join(c if c in "ACGT" else "A" for c in ad1)
        reads = [make_read(rng, plain if rng.random() < 0.5 else ad2, int(rng.integers(0, 6))) for _ in range(30)]
        reads += [mutate(rng, ad2, 0.05, 0.02, 0.02) + rnd(rng, 30) for _ in range(10)]  # 5' adapter at the start
        run_case(hs, cfg, [r for r in reads if r], stats)
    assert stats.get(("mode", 0), 0) > 100 and stats.get("matches", 0) > 20, stats


@pytest.mark.parametrize("seed", range(6))
def test_per_adapter_parameters_on_the_bit_parallel_search(hs, seed):
    """Two 3' adapters whose ;e= / ;o= differ from each other and from -e / -O: every adapter's own error table and
    minimum overlap must reach the bit-parallel search (they stay on the fast path) as they reach the literal DP."""
    rng = np.random.default_rng(4600 + seed)
    stats = {}
    for rep in range(5):
        ad1, ad2 = make_adapter(rng, "random"), make_adapter(rng, ["repeat", "two_blocks", "random"][rep % 3])
        spec1 = ad1 + ";e=%s" % rng.choice(["0", "0.05", "0.2", "0.25"]) + (";o=%d" % int(rng.integers(1, 9)) if rng.random() < 0.7 else "")
        spec2 = ad2 + (";o=%d" % int(rng.integers(1, 9))) + (";e=%s" % rng.choice(["0.1", "0.3"]) if rng.random() < 0.5 else "")
        cfg = P.TrimConfig(adapters=[("back", spec1), ("back", spec2)], error_rate=0.12, overlap=3, indels=True)
        reads = [make_read(rng, ad1 if rng.random() < 0.5 else ad2, int(rng.integers(0, 6))) for _ in range(60)]
        run_case(hs, cfg, [r for r in reads if r], stats)
    assert stats.get(("mode", 1), 0) > 100 and stats.get(("mode", 3), 0) > 50 and stats.get("matches", 0) > 100, stats


def test_more_adapters_than_the_bit_parallel_kernels_take(hs):
    """Seven adapters (a kit's worth through file:): fill_dev_params sends them to the full-DP search, and
    AdapterCutter's choice among them (most matches, then fewest errors, then the first) is the oracle's."""
    rng = np.random.default_rng(4800)
    ads = [make_adapter(rng, "random") for _ in range(5)] + [make_adapter(rng, "repeat"), make_adapter(rng, "two_blocks")]
    cfg = P.TrimConfig(adapters=[("back", a) for a in ads], error_rate=0.15, overlap=4, times=2, quality_cutoff=None)
    cp, pp = P.build_trim_params(cfg), py_params(cfg)
    fast_ok = C.c_int(1)
    err = C.create_string_buffer(512)
    assert hs.hs_set_params(C.byref(cp), C.byref(fast_ok), err, 512) == 0, err.value
    assert cp.n_adapters == 7 and fast_ok.value == 0
    hs.hs_pipeline.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int32)]
    win = (C.c_int32 * (2 * cp.n_mods))()
    mods = pp.modifiers()
    used = set()
    for it in range(1500):
        a = ads[int(rng.integers(len(ads)))]
        seq = make_read(rng, a, int(rng.integers(0, 5)))
        if rng.random() < 0.3:
            seq += mutate(rng, ads[int(rng.integers(len(ads)))], 0.05, 0.02, 0.02)
        if not seq:
            continue
        qual = "I" * len(seq)
        hs.hs_pipeline(seq.encode(), qual.encode(), len(seq), win)
        start, stop = 0, len(seq)
        for mi, mod in enumerate(mods):
            start, stop = po.apply_modifier(mod, seq, qual, start, stop, pp)
            assert (win[2 * mi], win[2 * mi + 1]) == (start, stop), (seq, mi, (win[2 * mi], win[2 * mi + 1]), (start, stop))
        which, mt = po.best_match(pp.adapters, seq, pp.compat)
        if mt is not None:
            used.add(pp.adapters.index(which))
    assert len(used) == 7, used


WHERE_SPECS = {  # cutadapt's notation of every placement (params.parse_adapter_spec)
    "back": ("back", "%s"), "front": ("front", "%s"), "suffix": ("back", "%s$"), "prefix": ("front", "^%s"),
    "back_not_internal": ("back", "%sX"), "front_not_internal": ("front", "X%s"),
}


def placed_read(rng, ad, where):
    """A read that carries (a damaged / partial copy of) the adapter where the placement looks for it -- or elsewhere."""
    r = rng.random()
    body = mutate(rng, ad, 0.06, 0.03, 0.03) if r < 0.7 else ad
    if rng.random() < 0.3:  # partial adapter: its start (3' forms) or its end (5' forms)
        cut = int(rng.integers(1, len(ad) + 1))
        body = body[:cut] if where in ("back", "suffix", "back_not_internal") else body[-cut:]
    ins = rnd(rng, int(rng.integers(0, 40)))
    pos = rng.random()
    if pos < 0.45:  # at the end the placement anchors
        s = ins + body if where in ("back", "suffix", "back_not_internal") else body + ins
    elif pos < 0.7:  # at the other end
        s = body + ins if where in ("back", "suffix", "back_not_internal") else ins + body
    elif pos < 0.9:  # inside
        s = ins + body + rnd(rng, int(rng.integers(1, 20)))
    else:
        s = rnd(rng, int(rng.integers(1, 60)))
    if s and rng.random() < 0.3:
        s = list(s)
        for _ in range(int(rng.integers(1, 4))):
            j = int(rng.integers(len(s)))
            s[j] = str(rng.choice(np.array(list("NNNRYKMnacgt"))))
        s = "".join(s)
    if rng.random() < 0.15:  # a copy that only matches through the read's wildcards, then a literal one
        dmg = list(ad)
        dmg[int(rng.integers(len(dmg)))] = "N"
        s = rnd(rng, int(rng.integers(0, 12))) + "".join(dmg) + rnd(rng, int(rng.integers(0, 8))) + ad + rnd(rng, int(rng.integers(0, 6)))
    return s


@pytest.mark.parametrize("seed", range(12))
def test_every_placement_and_read_wildcards_equal_the_oracle(hs, seed):
    """locate<MAXM, true>: anchored (^SEQ, SEQ$) and non-internal (XSEQ, SEQX) adapters, --match-read-wildcards, both
    objectives, with and without indels, against pyoracle.locate."""
    rng = np.random.default_rng(4400 + seed)
    stats = {}
    wheres = sorted(WHERE_SPECS)
    for rep in range(12):
        where = wheres[(seed + rep) % len(wheres)]
        kind, fmt = WHERE_SPECS[where]
        ad = make_adapter(rng, ["random", "repeat", "two_blocks", "prefix_repeat", "homopolymer"][rep % 5])
        if rng.random() < 0.3:
            ad = list(ad)
            ad[int(rng.integers(len(ad)))] = "N"
            ad = "".join(ad)
            if set(ad) == {"N"}:
                ad = "A" + ad
        cfg = P.TrimConfig(adapters=[(kind, fmt % ad)], error_rate=float(rng.choice([0.0, 0.1, 0.15, 0.25])),
                           overlap=int(rng.integers(1, 7)), indels=bool(rng.random() < 0.7),
                           match_read_wildcards=bool(rng.random() < 0.4), cutadapt_compat="4" if rng.random() < 0.25 else "2-3")
        plain = "".join(c if c in "ACGT" else "C" for c in ad)
        reads = [placed_read(rng, plain, where) for _ in range(50)]
        cp = P.build_trim_params(cfg)
        pp = py_params(cfg)
        assert pp.adapters[0].where == where and cp.adapters[0].where == abi.WHERE[where]
        fast_ok = C.c_int(0)
        err = C.create_string_buffer(512)
        assert hs.hs_set_params(C.byref(cp), C.byref(fast_ok), err, 512) == 0, err.value
        assert fast_ok.value == (1 if where == "back" and cfg.indels and not cfg.match_read_wildcards and cfg.cutadapt_compat != "4" else 0)
        out = (C.c_int32 * 4)()
        for read in reads:
            if not read:
                continue
            raw = read.encode()
            # Adapter.match_to: the literal comparison (find / startswith / endswith) first, else the alignment
            exp = po.match_to(pp.adapters[0], read, pp.compat)
            rc = hs.hs_locate(0, 0, raw, len(raw), 0, len(raw), out)
            got = tuple(out) if rc == 1 else None
            exp4 = None if exp is None else (exp[2], exp[3], exp[4], exp[5])
            assert got == exp4, (cfg.adapters, cfg.error_rate, cfg.overlap, cfg.indels, cfg.match_read_wildcards, cfg.cutadapt_compat, read, got, exp4)
            # ... and the shortcut only ever differs from the alignment through the read's wildcards (an occurrence through
            # an N in front of the literal one)
            al = po.locate(pp.adapters[0], read, pp.compat)
            if not cfg.match_read_wildcards:
                assert (al is None) == (exp is None) and (al is None or (al[2], al[3], al[5]) == (exp[2], exp[3], exp[5])), (cfg.adapters, read, al, exp)
            elif al != exp:
                stats["literal_first"] = stats.get("literal_first", 0) + 1
            stats[where] = stats.get(where, 0) + (exp is not None)
    assert sum(v for k, v in stats.items() if k != "literal_first") > 60, stats
    if seed == 0:
        print(stats)


PIPELINES = [
    dict(adapters=[("back", "^TTAGGC...TGGAATTCTCGGGTGCC")]),                                # the documented -a form: anchored 5' half, optional 3' half
    dict(adapters=[("back", "TTAGGC...TGGAATTCTCGGGTGCC")], times=2),                        # neither half required
    dict(adapters=[("front", "TTAGGC;optional...TGGAATTCTCGGGTGCC;e=0.2")], quality_cutoff="15"),
    dict(adapters=[("front", "^TTAGGC...TGGAATTCTCGGGTGCC$")], indels=False),
    dict(adapters=[("front", "^GTTCAGAGTTC"), ("back", "TGGAATTCTCGGGTGCCX")], times=2, trim_n=True, cut=[1, -1]),
    dict(adapters=[("back", "TGGAATTCTCGGGTGCC$;e=0.2"), ("front", "XGTTCAGAGTTC;o=5")], match_read_wildcards=True, nextseq_trim=20),
    dict(adapters=[("back", "TGGAATTCNNGGGTGCC;noindels"), ("back", "AAAAAAAA$")], match_read_wildcards=True, cutadapt_compat="4"),
]


@pytest.mark.parametrize("case", range(len(PIPELINES)))
def test_modifier_pipeline_with_linked_anchored_and_parameterised_adapters(hs, case):
    """apply_mod / best_match of the full-DP path (linked pairs with optional halves, every placement, per-adapter
    parameters, times > 1) against the Python oracle's modifier pipeline, window by window."""
    cfg = P.TrimConfig(**PIPELINES[case])
    cp, pp = P.build_trim_params(cfg), py_params(cfg)
    fast_ok = C.c_int(0)
    err = C.create_string_buffer(512)
    assert hs.hs_set_params(C.byref(cp), C.byref(fast_ok), err, 512) == 0, err.value
    hs.hs_pipeline.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int32)]
    rng = np.random.default_rng(4500 + case)
    mods = pp.modifiers()
    assert len(mods) == cp.n_mods
    win = (C.c_int32 * (2 * cp.n_mods))()
    front5, back3 = "TTAGGC", "TGGAATTCTCGGGTGCC"
    changed = 0
    for it in range(600):
        f = mutate(rng, front5 if case < 4 else "GTTCAGAGTTC", 0.05, 0.02, 0.02) if rng.random() < 0.7 else ""
        if f and rng.random() < 0.2:
            f = rnd(rng, int(rng.integers(1, 4))) + f  # not at the very start: an anchored 5' half misses it
        b = mutate(rng, back3, 0.05, 0.02, 0.02) if rng.random() < 0.7 else ""
        if b and rng.random() < 0.4:
            b = b[: int(rng.integers(1, len(b) + 1))]
        elif b and rng.random() < 0.5:
            b += rnd(rng, int(rng.integers(1, 12)))
        seq = f + rnd(rng, int(rng.integers(0, 35))) + b
        if rng.random() < 0.15:  # a copy of the 3' adapter that matches through an N only, then a literal one (match_to's shortcut)
            dmg = list(back3)
            dmg[int(rng.integers(len(dmg)))] = "N"
            seq = f + rnd(rng, int(rng.integers(0, 25))) + "".join(dmg) + rnd(rng, int(rng.integers(0, 6))) + back3 + rnd(rng, int(rng.integers(0, 5)))
        if not seq:
            continue
        if rng.random() < 0.2:
            seq = list(seq)
            seq[int(rng.integers(len(seq)))] = str(rng.choice(np.array(list("NNRn"))))
            seq = "".join(seq)
        q = np.clip(38 - (np.arange(len(seq)) * int(rng.integers(0, 40)) // len(seq)) + rng.integers(-5, 6, len(seq)), 2, 41)
        qual = "".join(chr(33 + int(v)) for v in q)
        assert hs.hs_pipeline(seq.encode(), qual.encode(), len(seq), win) == len(mods)
        start, stop = 0, len(seq)
        for mi, mod in enumerate(mods):
            start, stop = po.apply_modifier(mod, seq, qual, start, stop, pp)
            assert (win[2 * mi], win[2 * mi + 1]) == (start, stop), (PIPELINES[case], seq, qual, mi, (win[2 * mi], win[2 * mi + 1]), (start, stop))
        changed += (start, stop) != (0, len(seq))
    assert changed > 150, changed


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_modifier_pipeline_of_every_named_configuration(hs, name):
    """The configurations the GPU suite runs (tests/util.py CONFIGS), through the kernels' full-DP modifier pipeline
    compiled for the host, window by window against the Python oracle: what the GPU tests establish on a device for the
    generic kernel, established for its per-read code on every CPU run."""
    cfg = CONFIGS[name]
    cp, pp = P.build_trim_params(cfg), py_params(cfg)
    fast_ok = C.c_int(0)
    err = C.create_string_buffer(512)
    assert hs.hs_set_params(C.byref(cp), C.byref(fast_ok), err, 512) == 0, err.value
    hs.hs_pipeline.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int32)]
    mods = pp.modifiers()
    assert len(mods) == cp.n_mods
    win = (C.c_int32 * (2 * max(cp.n_mods, 1)))()
    data = random_fastq(1500, seed=zlib.crc32(name.encode()) % 1000 + 3, n_rate=0.02, lower_rate=0.01, **CONFIG_DATA.get(name, {}))
    changed = 0
    for _nm, seq, qual in po.parse_fastq(data):
        if not seq:
            continue
        assert hs.hs_pipeline(seq.encode(), qual.encode(), len(seq), win) == len(mods)
        start, stop = 0, len(seq)
        for mi, mod in enumerate(mods):
            start, stop = po.apply_modifier(mod, seq, qual, start, stop, pp)
            assert (win[2 * mi], win[2 * mi + 1]) == (start, stop), (name, seq, qual, mi, (win[2 * mi], win[2 * mi + 1]), (start, stop))
        changed += (start, stop) != (0, len(seq))
    assert changed > 300, changed


def test_quality_scans_equal_the_oracle(hs):
    """nextseq_trim_index / quality_trim_index of the kernels on random and adversarial
    quality strings: ties of the running maximum, scans that never go negative, empty reads, both Phred offsets."""
    hs.hs_nextseq.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    hs.hs_quality.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(77)
    n = 0
    for it in range(6000):
        L = int(rng.integers(0, 120))
        base = 33 if rng.random() < 0.8 else 64
        mode = it % 5
        if mode == 0:
            q = rng.integers(0, 42, L)
        elif mode == 1:  # decaying qualities
            q = np.clip(38 - (np.arange(L) * rng.integers(0, 50) // max(L, 1)) + rng.integers(-5, 6, L), 0, 41)
        elif mode == 2:  # two levels around the cut-off: many ties
            q = rng.choice([19, 20, 21], L)
        elif mode == 3:
            q = np.full(L, int(rng.integers(0, 42)))
        else:
            q = rng.integers(0, 94 - (base - 33), L)  # any printable byte
        qual = "".join(chr(base + int(v)) for v in q)
        seq = "".join(rng.choice(np.array(list("ACGTGGN")), L))
        cutoff = int(rng.integers(0, 41))
        q5 = int(rng.integers(0, 41)) if rng.random() < 0.5 else 0
        assert hs.hs_nextseq(seq.encode(), qual.encode(), L, cutoff, base) == po.nextseq_trim_index(seq, qual, cutoff, base), (seq, qual, cutoff, base)
        s, e = C.c_int(-1), C.c_int(-1)
        hs.hs_quality(qual.encode(), L, q5, cutoff, base, C.byref(s), C.byref(e))
        assert (s.value, e.value) == tuple(po.quality_trim_index(qual, q5, cutoff, base)), (qual, q5, cutoff, base)
        n += 1
    assert n == 6000
