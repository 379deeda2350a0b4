"""Shared helpers for the tests: small seeded FASTQ generator and parameter conversion."""
import argparse

import numpy as np

import mirge_b200  # noqa: F401
from mirge_b200 import params as P
from oracle import pyoracle as po

ILL = "TGGAATTCTCGGGTGCCAAGGAACTCCAG"
QIA = "AACTGTAGGCACCATCAAT"
LONG_AD = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACATCACGATCTCGTATGCC"


def py_adapters(cfg: P.TrimConfig):
    """The adapters of a configuration as the Python oracle's objects (specification language: params.parse_adapter_specs)."""
    ads = []
    for k, s in cfg.adapters:
        for sp in P.parse_adapter_specs(k, s):
            mk = lambda where, seq, prm: po.Adapter(where, seq, prm.get("max_error_rate", cfg.error_rate), prm.get("min_overlap", cfg.overlap),
                                                    prm.get("indels", cfg.indels), cfg.match_adapter_wildcards, cfg.match_read_wildcards)
            if sp.where == "linked":
                ads.append(po.LinkedAdapter(mk(sp.where5, sp.sequence, sp.params), mk(sp.where2, sp.sequence2, sp.params2),
                                            sp.front_required, sp.back_required))
            else:
                ads.append(mk(sp.where, sp.sequence, sp.params))
    return ads


def py_params(cfg: P.TrimConfig) -> po.TrimParams:
    ads = py_adapters(cfg)
    return po.TrimParams(adapters=ads, times=cfg.times, nextseq_trim=cfg.nextseq_trim,
                         quality_cutoff=cfg.quality_cutoff, quality_base=cfg.quality_base, trim_n=cfg.trim_n,
                         cut=list(cfg.cut), minimum_length=cfg.minimum_length, umi=cfg.umi(),
                         qiagenumi=cfg.qiagenumi, count_mode=cfg.count_mode, compat=cfg.cutadapt_compat or "2-3")


def random_fastq(n, seed=0, L=50, adapter=ILL, umi3=0, n_rate=0.01, lower_rate=0.0, err=0.03, indel=0.01,
                 pool=200, crlf=False, final_newline=True, varlen=False, front=None):
    """Small-RNA-like reads: insert (from a pool, so keys repeat) + adapter (with errors) + tail."""
    rng = np.random.default_rng(seed)
    B = np.array(list("ACGT"))
    inserts = ["".join(rng.choice(B, rng.integers(10, 40))) for _ in range(pool)]
    out = []
    for i in range(n):
        if rng.random() < 0.9:
            ins = inserts[int(rng.zipf(1.3)) % pool]
        else:
            ins = "".join(rng.choice(B, rng.integers(0, 45)))
        ad = list(adapter)
        j = 0
        while j < len(ad):
            r = rng.random()
            if r < err:
                ad[j] = str(rng.choice(B))
            elif r < err + indel / 2:
                del ad[j]
                continue
            elif r < err + indel:
                ad.insert(j, str(rng.choice(B)))
                j += 1
            j += 1
        umi = "".join(rng.choice(B, umi3)) if umi3 else ""
        s = (front or "") + ins + ("".join(ad) if rng.random() < 0.9 else "") + umi + "".join(rng.choice(B, L))
        ln = L if not varlen else int(rng.integers(0, L + 1))
        s = list(s[:ln])
        for j in range(len(s)):
            r = rng.random()
            if r < n_rate:
                s[j] = "N"
            elif r < n_rate + lower_rate:
                s[j] = s[j].lower()
        s = "".join(s)
        q0, q1 = rng.integers(25, 41), rng.integers(2, 30)
        q = np.clip(np.linspace(q0, q1, max(len(s), 1))[: len(s)] + rng.integers(-5, 6, len(s)), 0, 41).astype(int)
        if rng.random() < 0.1 and len(s) > 12:  # NextSeq dark-cycle poly-G tail with high quality
            t = int(rng.integers(1, 12))
            s = s[:-t] + "G" * t
            q[-t:] = 35
        nl = "\r\n" if crlf else "\n"
        out.append("@SYN.%d %d length=%d%s%s%s+%s%s%s"
                   % (i, i, len(s), nl, s, nl, nl, "".join(chr(33 + int(v)) for v in q), nl))
    data = "".join(out)
    if not final_newline:
        data = data.rstrip("\r\n")
    return data.encode()


CONFIGS = {
    "default": P.TrimConfig(adapters=[("back", ILL)]),
    "release": P.TrimConfig(adapters=[("back", ILL)], count_mode="release"),
    "nextseq": P.TrimConfig(adapters=[("back", ILL)], nextseq_trim=20, quality_cutoff="20"),
    "q5q3_nx_cut": P.TrimConfig(adapters=[("back", ILL)], quality_cutoff="15,20", trim_n=True, cut=[2, -1]),
    "umi44": P.TrimConfig(adapters=[("back", ILL)], uniq_mol_ids="4,4"),
    "umi40": P.TrimConfig(adapters=[("back", ILL)], uniq_mol_ids="4,0", count_mode="release"),
    "qiagen": P.TrimConfig(adapters=[("back", QIA)], uniq_mol_ids="0,12", qiagenumi=True),
    "front_back": P.TrimConfig(adapters=[("back", ILL), ("front", "GTTCAGAGTTCTACAGTCCGACGATC")]),
    "noindel": P.TrimConfig(adapters=[("back", ILL)], indels=False),
    "wild": P.TrimConfig(adapters=[("back", "TGGAATTCNNGGGTGCCAAGGRACTCCAG")], error_rate=0.2, overlap=5),
    "noq_m1": P.TrimConfig(adapters=[("back", ILL)], quality_cutoff=None, minimum_length=1, times=2),
    "long_adapter": P.TrimConfig(adapters=[("back", LONG_AD)]),
    "reads150": P.TrimConfig(adapters=[("back", ILL)], nextseq_trim=20, quality_cutoff="20", trim_n=True, cut=[1]),
    "reads200": P.TrimConfig(adapters=[("back", ILL)], quality_cutoff="20"),
    # linked adapters -g "A...B" (quick_start.md:208-220): both halves required, 3' half searched behind the 5' match
    "linked": P.TrimConfig(adapters=[("front", "TTAGGC...TGGAATTCTCGGGTGCCAAGGAACTCCAGT")]),
    "linked_x2": P.TrimConfig(adapters=[("front", "CAGTCCGACGATC..." + ILL), ("back", "AAAAAAAAAA")], times=2, indels=False,
                              uniq_mol_ids="2,2"),
    # cutadapt >= 4 objective (score instead of matches): full-DP kernel
    "compat4": P.TrimConfig(adapters=[("back", ILL)], cutadapt_compat="4"),
    "compat4_fb": P.TrimConfig(adapters=[("back", ILL), ("front", "GTTCAGAGTTCTACAGTCCGACGATC")], cutadapt_compat="4", times=2),
}

# keyword arguments for random_fastq that exercise each configuration
CONFIG_DATA = {
    "qiagen": dict(adapter=QIA, umi3=12, L=75),
    "umi44": dict(L=60),
    "umi40": dict(L=60),
    "front_back": dict(front="CAGTCCGACGATC"),
    "wild": dict(adapter=ILL),
    "long_adapter": dict(adapter=LONG_AD, L=90),
    "noq_m1": dict(varlen=True),
    "reads150": dict(L=150),  # still inside the packed-read fast path (PACK_WORDS * 16 = 160 bases)
    "reads200": dict(L=200),  # beyond it: whole-pipeline kernel per read
    "linked": dict(front="TTAGGC", err=0.05),
    "linked_x2": dict(front="CAGTCCGACGATC", L=70),
    "compat4": dict(err=0.06, indel=0.04),
    "compat4_fb": dict(front="CAGTCCGACGATC", err=0.06, indel=0.04),
}


def make_args(**kw):
    a = argparse.Namespace(
        adapters=[("back", ILL)], error_rate=0.12, overlap=3, indels=True, match_adapter_wildcards=True,
        match_read_wildcards=False, times=1, action="trim", nextseq_trim=None, quality_cutoff="10", phred64=33,
        trim_n=False, cut=[], minimum_length=16, uniq_mol_ids=None, qiagenumi=False, umiDedup=False, quiet=True,
        tcf_out=False, bam_out=False, tRNA_frag=False, spikeIn=False, threads=1, organism_name="human",
        libraries_path=None, bowtie_path=None, bowtieVersion="1.3.0", cutadaptVersion=(3, 1), buffer_size=4000000, fasta=False)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def write_ebwt(basename, names, seqs, line_rate=6, lines_per_side=1, ftab_chars=4, bwt_pad=0):
    """Minimal bowtie 1 index writer for the loader tests: real ``.3.ebwt`` / ``.4.ebwt`` and a ``.1.ebwt`` whose
    header, table sizes and name block follow the published layout (the BWT region itself is zero bytes of the
    right size; ``bwt_pad`` adds bytes there to emulate a layout the header arithmetic does not predict)."""
    import struct

    recs, bases = [], []
    for s in seqs:
        s = s.upper()
        first, i, n = 1, 0, len(s)
        while i < n:
            j = i
            while j < n and s[j] not in "ACGT":
                j += 1
            k = j
            while k < n and s[k] in "ACGT":
                k += 1
            if k > j:
                recs.append((j - i, k - j, first))
                first = 0
                bases.append(s[j:k])
            i = k
        if first:  # reference without any unambiguous base: an empty first record keeps the name aligned
            recs.append((0, 0, 1))
    text = "".join(bases)
    total = len(text)
    with open(basename + ".3.ebwt", "wb") as f:
        f.write(struct.pack("<iI", 1, len(recs)))
        for off, ln, first in recs:
            f.write(struct.pack("<IIB", off, ln, first))
    codes = np.array(["ACGT".index(c) for c in text] + [0] * ((-total) % 4), dtype=np.uint8).reshape(-1, 4)
    packed = (codes[:, 0] | (codes[:, 1] << 2) | (codes[:, 2] << 4) | (codes[:, 3] << 6)).astype(np.uint8)
    packed.tofile(basename + ".4.ebwt")
    side_sz = (1 << line_rate) * lines_per_side
    side_bwt_sz = side_sz - 8
    bwt_sz = total // 4 + 1
    n_side_pairs = (bwt_sz + 2 * side_bwt_sz - 1) // (2 * side_bwt_sz)
    with open(basename + ".1.ebwt", "wb") as f:
        f.write(struct.pack("<7i", 1, total, line_rate, lines_per_side, 5, ftab_chars, 0))
        f.write(struct.pack("<I", len(seqs)))
        f.write(struct.pack("<%dI" % len(seqs), *[len(s) for s in seqs]))
        f.write(struct.pack("<I", len(recs)))
        f.write(b"\x00" * (12 * len(recs)))
        f.write(b"\x00" * (n_side_pairs * 2 * side_sz + bwt_pad))
        f.write(struct.pack("<I", 0) + struct.pack("<5I", 0, 1, 2, 3, total))
        f.write(b"\x00" * (4 * ((1 << (2 * ftab_chars)) + 1)))
        f.write(b"\x00" * (4 * 2 * ftab_chars))
        f.write(("\n".join(names) + "\n").encode() + b"\x00")


def host_harness_flags():
    """g++ flags for the device headers compiled for the host (tests/test_*_host.py).  MIRGE_TEST_SANITIZE=1 adds
    AddressSanitizer + UBSan, for which the interpreter must run with the sanitizer runtimes preloaded:
        LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \\
        MIRGE_TEST_SANITIZE=1 python -m pytest tests/test_adapter_search_host.py tests/test_annotate_verify_host.py tests/test_key_format_host.py
    (the kernels' per-read arithmetic then runs with every out-of-bounds access of a local array or staging buffer and every
    undefined shift / overflow turned into a failure)."""
    import os

    if os.environ.get("MIRGE_TEST_SANITIZE") == "1":
        return ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]
    return ["-O2"]
