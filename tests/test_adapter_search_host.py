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

import numpy as np
import pytest

import mirge_b200
from mirge_b200 import abi
from mirge_b200 import params as P
from oracle import pyoracle as po
from tests.util import py_params

HERE = os.path.dirname(os.path.abspath(__file__))
B = np.array(list("ACGT"))


@pytest.fixture(scope="module")
def hs(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hs") / "libadapter_search_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(mirge_b200.PACKAGE_DIR, "csrc"),
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
