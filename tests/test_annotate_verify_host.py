"""The per-sequence arithmetic of the annotation kernels (mirge3.0_b200/csrc/annotate_verify.cuh: round window with the
poly-T rule and -5/-3 trimming, 2-bit query words with always-mismatch masks, alignment verification under the -n / -v
policies with reference bounds and ambiguous reference bases, division-free seed piece boundaries, k-mer extraction)
compiled for the host and held against oracle/pyoracle.py (hits + canonical pick) on small libraries
built to provoke the rules: references with N, near-duplicates one and two mismatches apart, matches at reference
ends, queries with N and lower case, T tails, lengths around the 25/26 and 28-base limits."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from tests.util import host_harness_flags
import mirge_b200
from mirge_b200 import abi
from mirge_b200 import libraries as LB
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
B = np.array(list("ACGT"))
NO_HIT = 0xFFFFFFFFFFFFFFFF


@pytest.fixture(scope="module")
def hv(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hv") / "libannotate_verify_host.so")
    subprocess.check_call(["g++"] + host_harness_flags() + ["-std=c++17", "-shared", "-fPIC", "-I", os.path.join(mirge_b200.PACKAGE_DIR, "csrc"),
                           "-o", so, os.path.join(HERE, "annotate_verify_harness.cpp")])
    lib = C.CDLL(so)
    lib.hv_best_hit.restype = C.c_uint64
    lib.hv_best_hit.argtypes = [C.POINTER(abi.Library), C.POINTER(abi.RoundPolicy), C.c_void_p, C.POINTER(C.c_int), C.c_void_p,
                                C.POINTER(C.c_int)]
    lib.hv_pieces.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


class HostLibrary:
    """The arrays of a mirge_library the verification needs, built on the host the way DeviceLibrary builds them."""

    def __init__(self, seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.int64)
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        text = torch.frombuffer(bytearray("".join(seqs).encode()), dtype=torch.uint8)
        packed, nmask = LB.pack_text(text)
        self.packed = np.concatenate([packed.numpy().view(np.uint32), np.zeros(abi.LIB_PAD_WORDS, dtype=np.uint32)])
        self.nmask = nmask.numpy().view(np.uint32).copy()
        self.ref_off = off.astype(np.uint32)
        shift = 6
        starts = np.arange((int(off[-1]) >> shift) + 1, dtype=np.int64) << shift
        self.ref_block = np.clip(np.searchsorted(off, starts, side="right") - 1, 0, max(len(seqs) - 1, 0)).astype(np.uint32)
        p = lambda a: a.ctypes.data
        self.struct = abi.Library(p(self.packed), p(self.nmask), p(self.ref_off), len(seqs), int(off[-1]), 0, 0, 0, 4, 0,
                                  p(self.ref_block), shift, 0, 0)


def pack_key(s: str) -> np.ndarray:
    """[len | n_exc << 16][2-bit payload][exceptions pos << 8 | byte]: DESIGN.md section 3."""
    n = len(s)
    pay = np.zeros((n + 15) // 16, dtype=np.uint32)
    exc = []
    for i, ch in enumerate(s):
        code = "ACGT".find(ch)
        if code < 0:
            exc.append((i << 8) | ord(ch))
        else:
            pay[i >> 4] |= np.uint32(code << (2 * (i & 15)))
    return np.concatenate([np.array([n | (len(exc) << 16)], dtype=np.uint32), pay, np.array(exc, dtype=np.uint32)])


def rnd(rng, n):
    return "".join(rng.choice(B, n))


def sub(rng, s, k):
    s = list(s)
    for j in rng.choice(len(s), size=min(k, len(s)), replace=False):
        s[j] = str(rng.choice([c for c in "ACGT" if c != s[j]]))
    return "".join(s)


def make_case(rng):
    refs = [rnd(rng, int(rng.integers(18, 90))) for _ in range(int(rng.integers(3, 9)))]
    refs.append(sub(rng, refs[0], 1))          # near-duplicates: the canonical pick must take the lower index / offset
    refs.append(sub(rng, refs[1], 2))
    refs.append(refs[2][:10] + "N" + refs[2][11:]) if len(refs[2]) > 12 else None
    refs.append(refs[0][5:] + refs[0][:5])
    refs = [r for r in refs if r]
    queries = []
    for _ in range(40):
        r = refs[int(rng.integers(len(refs)))].replace("N", "A")
        L = int(rng.choice([16, 20, 22, 24, 25, 26, 27, 28, 29, 30, 33, 40, 60]))
        L = min(L, len(r))
        o = int(rng.integers(0, len(r) - L + 1)) if rng.random() < 0.7 else (0 if rng.random() < 0.5 else len(r) - L)
        q = r[o : o + L]
        k = int(rng.choice([0, 0, 1, 1, 2, 3]))
        q = sub(rng, q, k)
        m = rng.random()
        if m < 0.15:
            q = q + "T" * int(rng.integers(1, 6))       # poly-T tails (round 3 strips T{3,}$)
        elif m < 0.25:
            q = str(rng.choice(B)) + q + rnd(rng, 2)      # isomiR-style ends (round 8 trims 1 and 2)
        elif m < 0.32 and len(q) > 3:
            j = int(rng.integers(len(q)))
            q = q[:j] + ("N" if rng.random() < 0.5 else q[j].lower()) + q[j + 1:]
        queries.append(q)
    queries += [rnd(rng, int(rng.integers(1, 40))) for _ in range(6)] + ["TTT", "TTTT", "ACGTTTT", "A"]
    return refs, queries


@pytest.mark.parametrize("seed", range(10))
def test_verification_scan_equals_the_oracle(hv, seed):
    rng = np.random.default_rng(6100 + seed)
    refs, queries = make_case(rng)
    lib = HostLibrary(refs)
    olib = po.Library(["r%d" % i for i in range(len(refs))], refs)
    pols = LB.round_policies()
    searched = C.c_int(0)
    qlen = C.c_int(0)
    qwords = np.zeros(2 * 40, dtype=np.uint32)
    n_hits = 0
    for q in queries:
        key = pack_key(q)
        for rnd_i in range(10):
            got = hv.hv_best_hit(C.byref(lib.struct), C.byref(pols[rnd_i]), key.ctypes.data, C.byref(searched), qwords.ctypes.data,
                                 C.byref(qlen))
            oq = po.round_query(q, rnd_i)  # None / "" when the round does not search this sequence's text
            # the length selection of rounds 0 / 1 is the kernel's business (select); the window rules are tested here
            if rnd_i in (0, 1):
                oq = q
            if not oq:
                assert searched.value == 0 or qlen.value == 0 or got == NO_HIT, (q, rnd_i)
                continue
            assert searched.value == 1 and qlen.value == len(oq), (q, rnd_i, oq, qlen.value)
            hs = po.hits(oq, olib, po.ROUND_POLICIES[rnd_i])
            if hs:
                mm, r, off = min((h[0], h[1], h[2]) for h in hs)
                exp = (mm << 56) | (r << 28) | off
                n_hits += 1
            else:
                exp = NO_HIT
            assert got == exp, (q, rnd_i, oq, hex(got), hex(exp))
    assert n_hits > 50


def test_seed_pieces_and_kmers(hv):
    """piece_bound == pi * R // np for every seed length the policies can produce; query_kmer16 == the first 16 bases of
    the piece, first base most significant (the order of the sorted index)."""
    rng = np.random.default_rng(5)
    lib = hv
    out = np.zeros(16, dtype=np.uint32)
    for L in list(range(1, 130)) + [511, 512]:
        s = rnd(rng, L)
        qw = np.zeros((L + 15) // 16 + 1, dtype=np.uint32)
        for i, ch in enumerate(s):
            qw[i >> 4] |= np.uint32("ACGT".find(ch) << (2 * (i & 15)))
        for seed_len in (0, 28):
            for seed_mm in (0, 1, 2, 3):
                np_ = lib.hv_pieces(qw.ctypes.data, L, seed_len, seed_mm, out.ctypes.data)
                R = L if seed_len == 0 else min(seed_len, L)
                assert np_ == seed_mm + 1
                for pi in range(np_ + 1):
                    assert int(out[2 * pi]) == pi * R // np_, (L, seed_len, seed_mm, pi)
                for pi in range(np_):
                    a = pi * R // np_
                    k = 0
                    for j in range(16):
                        code = "ACGT".find(s[a + j]) if a + j < L else 0
                        k |= code << (2 * (15 - j))
                    assert int(out[2 * pi + 1]) == k, (L, seed_len, seed_mm, pi)
