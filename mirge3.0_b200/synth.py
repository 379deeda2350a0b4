"""Synthetic benchmark inputs: seeded libraries with human-miRBase-like sizes
and small-RNA reads drawn from them (isomiR ends, substitutions, ligated adapters with errors, UMIs,
decaying qualities, NextSeq poly-G tails).  Libraries use ``numpy.random.default_rng(1000 + lib_id)``
as specified; reads are generated with torch on the target device (``torch.Generator`` seeded with
2000 + config) so that tens of millions of reads are produced in seconds on the GPU -- the bytes are
data, every parity check runs the oracle on the very same bytes."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

ILLUMINA = "TGGAATTCTCGGGTGCCAAGGAACTCCAG"
QIA_INNER = "AACTGTAGGCACCATCAAT"
QIA_OUTER = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACATCACGATCTCGTATGCCGTCTTCTGCTTG"

LIB_IDS = {"mirna": 0, "hairpin": 1, "mature_trna": 2, "pre_trna": 3, "snorna": 4, "rrna": 5, "ncrna_others": 6,
           "mrna": 7, "spike-in": 9}
# (count, min_len, max_len)
LIB_SPEC = {"mirna": (2656, 18, 25), "hairpin": (1917, 60, 120), "mature_trna": (432, 70, 90), "pre_trna": (600, 90, 150),
            "snorna": (1000, 60, 300), "rrna": (40, 120, 5000), "ncrna_others": (20000, 100, 1000),
            "mrna": (100000, 500, 4000), "spike-in": (52, 22, 22)}
PREFIX = {"mirna": "syn-miR-", "hairpin": "syn-mir-", "mature_trna": "syn-tRNA-", "pre_trna": "syn-pre-tRNA-",
          "snorna": "syn-snoRNA-", "rrna": "syn-rRNA-", "ncrna_others": "syn-ncRNA-", "mrna": "syn-mRNA-", "spike-in": "syn-spike-"}
ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class SynthLibrary:
    names: List[str]
    codes: np.ndarray  # uint8 0..3 (4 = N), concatenated
    off: np.ndarray  # int64 [n + 1]

    def seqs(self) -> List[bytes]:
        lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
        txt = lut[self.codes].tobytes()
        return [txt[int(self.off[i]) : int(self.off[i + 1])] for i in range(len(self.names))]

    def fasta(self) -> str:
        return "".join(">%s\n%s\n" % (n, s.decode()) for n, s in zip(self.names, self.seqs()))


@dataclass
class SynthLibraries:
    libs: Dict[str, SynthLibrary]
    mir_hairpin: np.ndarray  # for miRNA i: hairpin index or -1
    mir_offset: np.ndarray  # offset of miRNA i inside its hairpin

    def fasta_dict(self):
        return {k: (v.names, v.seqs()) for k, v in self.libs.items()}


def make_libraries(scale: float = 1.0, mrna_count: Optional[int] = None, n_rate: float = 0.001) -> SynthLibraries:
    """Seeded libraries; ``scale`` shrinks every entry count (tests), ``mrna_count`` overrides the mRNA
    library size (5 000 for config C1, 100 000 for C2/C5)."""
    libs: Dict[str, SynthLibrary] = {}
    for key, (cnt, lo, hi) in LIB_SPEC.items():
        rng = np.random.default_rng(1000 + LIB_IDS[key])
        n = max(4, int(round(cnt * scale)))
        if key == "mrna" and mrna_count is not None:
            n = int(mrna_count)
        lens = rng.integers(lo, hi + 1, size=n).astype(np.int64)
        if key == "mature_trna":
            lens += 3
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
        if key == "mature_trna":  # ...CCA 3' end
            for j, c in enumerate((1, 1, 0)):
                codes[off[1:] - 3 + j] = c
        if key in ("ncrna_others", "mrna") and n_rate > 0:
            npos = rng.integers(0, codes.size, size=int(codes.size * n_rate))
            codes[npos] = 4
        libs[key] = SynthLibrary([PREFIX[key] + str(i + 1) for i in range(n)], codes, off)
    # no duplicate sequences inside the short-entry libraries
    for key in ("mirna", "spike-in"):
        L = libs[key]
        s = L.seqs()
        assert len(set(s)) == len(s), "duplicate synthetic %s entry (change the seed)" % key
    # every hairpin embeds >= 1 miRNA at a random offset
    rng = np.random.default_rng(1000 + 1)
    mir, hp = libs["mirna"], libs["hairpin"]
    n_m, n_h = len(mir.names), len(hp.names)
    mir_h = np.full(n_m, -1, dtype=np.int64)
    mir_o = np.zeros(n_m, dtype=np.int64)
    for i in range(min(n_m, n_h)):
        ml = int(mir.off[i + 1] - mir.off[i])
        hl = int(hp.off[i + 1] - hp.off[i])
        o = int(rng.integers(3, hl - ml - 3))
        hp.codes[hp.off[i] + o : hp.off[i] + o + ml] = mir.codes[mir.off[i] : mir.off[i + 1]]
        mir_h[i], mir_o[i] = i, o
    return SynthLibraries(libs, mir_h, mir_o)


@dataclass
class ReadConfig:
    """Per-config read model."""

    cfg_id: int = 1
    L: int = 50
    adapter: str = ILLUMINA
    qiaseq: bool = False
    umi_len: int = 12
    polyg_frac: float = 0.0
    sample_seed: int = 0  # C4: per-sample abundance perturbation


CONFIGS = {
    1: ReadConfig(1, 50),
    2: ReadConfig(2, 75, polyg_frac=0.10),
    3: ReadConfig(3, 75, adapter=QIA_INNER, qiaseq=True),
    4: ReadConfig(4, 50),
    5: ReadConfig(5, 75),
}

CLASS_P = [0.70, 0.04, 0.06, 0.03, 0.05, 0.02, 0.04, 0.01, 0.05]
# class -> library key; tRNA class splits 3:1 into mature / pre (pre fragments get a TTTT tail)
CLASS_LIB = ["mirna", "hairpin", "mature_trna", "snorna", "rrna", "ncrna_others", "mrna", "spike-in", None]


class ReadGenerator:
    def __init__(self, libs: SynthLibraries, rc: ReadConfig, device="cpu"):
        self.libs, self.rc = libs, rc
        self.device = torch.device(device)
        dev = self.device
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(2000 + rc.cfg_id + 1000003 * rc.sample_seed)
        # one universe of 2-bit codes with per-library base offsets
        keys = ["mirna", "hairpin", "mature_trna", "pre_trna", "snorna", "rrna", "ncrna_others", "mrna", "spike-in"]
        self.base: Dict[str, int] = {}
        self.off: Dict[str, torch.Tensor] = {}
        self.cdf: Dict[str, torch.Tensor] = {}
        chunks = []
        pos = 0
        for k in keys:
            L = libs.libs[k]
            self.base[k] = pos
            chunks.append(torch.from_numpy(L.codes))
            self.off[k] = torch.from_numpy(L.off).to(dev) + pos
            n = len(L.names)
            # Zipf(1.1) abundance over a seeded permutation of the entries
            prng = np.random.default_rng(3000 + LIB_IDS[k] + 7919 * rc.sample_seed)
            w = 1.0 / np.arange(1, n + 1) ** 1.1
            w = w[prng.permutation(n)]
            if rc.sample_seed:
                w = w * prng.dirichlet(np.full(n, 2.0)) * n
            self.cdf[k] = torch.from_numpy(np.cumsum(w / w.sum())).to(dev)
            pos += L.codes.size
        self.universe = torch.cat(chunks).to(dev)
        self.mir_h = torch.from_numpy(libs.mir_hairpin).to(dev)
        self.mir_o = torch.from_numpy(libs.mir_offset).to(dev)
        ad = rc.adapter + (("N" * rc.umi_len) + QIA_OUTER if rc.qiaseq else "")
        self.ad_codes = torch.tensor(["ACGTN".index(c) for c in ad], dtype=torch.uint8, device=dev)
        self.class_cdf = torch.tensor(np.cumsum(CLASS_P), device=dev)

    # ---- helpers
    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.gen, device=self.device)

    def _randint(self, lo, hi, shape):
        return torch.randint(lo, hi, shape, generator=self.gen, device=self.device)

    def _pick(self, key: str, n: int) -> torch.Tensor:
        u = self._rand(n).to(torch.float64)
        return torch.searchsorted(self.cdf[key], u).clamp_(max=self.cdf[key].numel() - 1)

    def inserts(self, n: int):
        """(codes uint8 [n, 48] with 0..3, lens int64 [n]) of the biological inserts."""
        dev = self.device
        IM = 48
        cls = torch.searchsorted(self.class_cdf, self._rand(n).to(torch.float64)).clamp_(max=8)
        src0 = torch.zeros(n, dtype=torch.int64, device=dev)  # universe coordinate of insert base 0
        lo = torch.zeros(n, dtype=torch.int64, device=dev)  # template bounds (outside: random base)
        hi = torch.zeros(n, dtype=torch.int64, device=dev)
        flen = torch.zeros(n, dtype=torch.int64, device=dev)
        tail_t = torch.zeros(n, dtype=torch.int64, device=dev)  # appended T run (pre-tRNA fragments)
        for c in range(9):
            m = cls == c
            k = int(m.sum())
            if k == 0:
                continue
            if c == 8:  # random 16..40-mer
                flen[m] = self._randint(16, 41, (k,))
                continue
            key = CLASS_LIB[c]
            pre = None
            if c == 2:
                pre = self._rand(k) < 0.25
            e = self._pick(key, k)
            o0, o1 = self.off[key][e], self.off[key][e + 1]
            if c == 2:
                e2 = self._pick("pre_trna", k)
                p0, p1 = self.off["pre_trna"][e2], self.off["pre_trna"][e2 + 1]
                o0 = torch.where(pre, p0, o0)
                o1 = torch.where(pre, p1, o1)
            if c == 0:  # mature miRNA with isomiR ends, templated from its hairpin when it has one
                r5, r3 = self._rand(k), self._rand(k)
                s5 = torch.where(r5 < 0.04, -1, torch.where(r5 < 0.96, 0, 1))
                s3 = torch.where(r3 < 0.05, -2, torch.where(r3 < 0.20, -1, torch.where(r3 < 0.75, 0, torch.where(r3 < 0.90, 1, 2))))
                h = self.mir_h[e]
                has = h >= 0
                hs = self.off["hairpin"][h.clamp(min=0)]
                he = self.off["hairpin"][h.clamp(min=0) + 1]
                start = torch.where(has, hs + self.mir_o[e], o0)
                ml = o1 - o0
                src0[m] = start + s5
                flen[m] = ml - s5 + s3
                lo[m] = torch.where(has, hs, o0)
                hi[m] = torch.where(has, he, o1)
            elif c == 7:  # spike-in: whole entry
                src0[m], flen[m], lo[m], hi[m] = o0, o1 - o0, o0, o1
            else:
                fl = self._randint(18, 41, (k,))
                fl = torch.minimum(fl, o1 - o0)
                st = o0 + (self._rand(k) * (o1 - o0 - fl + 1).to(torch.float32)).to(torch.int64).clamp_(min=0)
                st = torch.minimum(st, o1 - fl)
                src0[m], flen[m], lo[m], hi[m] = st, fl, o0, o1
                if c == 2:
                    tail_t[m] = torch.where(pre, torch.full_like(fl, 4), torch.zeros_like(fl))
        j = torch.arange(IM, device=dev).unsqueeze(0)
        idx = src0.unsqueeze(1) + j
        inside = (idx >= lo.unsqueeze(1)) & (idx < hi.unsqueeze(1))
        codes = self.universe[idx.clamp(0, self.universe.numel() - 1)]
        rnd = self._randint(0, 4, (n, IM)).to(torch.uint8)
        codes = torch.where(inside & (codes < 4), codes, rnd)
        # sequencing substitutions 0.5 % / base
        sub = self._rand(n, IM) < 0.005
        codes = torch.where(sub, (codes + self._randint(1, 4, (n, IM)).to(torch.uint8)) % 4, codes)
        # pre-tRNA fragments end in TTTT
        tpos = (j >= flen.unsqueeze(1)) & (j < (flen + tail_t).unsqueeze(1))
        codes = torch.where(tpos, torch.full_like(codes, 3), codes)
        flen = (flen + tail_t).clamp_(max=IM)
        return codes, flen

    def block(self, n: int, first_index: int) -> torch.Tensor:
        """FASTQ bytes (uint8 tensor on self.device) of reads first_index .. first_index + n - 1."""
        dev, rc = self.device, self.rc
        L = rc.L
        ins, ilen = self.inserts(n)
        if rc.qiaseq:
            # PCR duplicates: a read repeats the previous molecule with probability 0.7 (Geometric(0.3) family size)
            dup = self._rand(n) < 0.7
            dup[0] = False
            src = torch.cummax(torch.where(dup, torch.zeros(n, dtype=torch.int64, device=dev), torch.arange(n, device=dev)), 0)[0]
            ins, ilen = ins[src], ilen[src]
        else:
            src = None
        IM = ins.shape[1]
        A = self.ad_codes.numel()
        ad = self.ad_codes.unsqueeze(0).expand(n, A).clone()
        if rc.qiaseq:
            umi = self._randint(0, 4, (n, rc.umi_len)).to(torch.uint8)
            umi = umi[src]
            a0 = len(rc.adapter)
            ad[:, a0 : a0 + rc.umi_len] = umi
            errmask = torch.ones(A, dtype=torch.bool, device=dev)
            errmask[a0 : a0 + rc.umi_len] = False
        else:
            errmask = torch.ones(A, dtype=torch.bool, device=dev)
        sub = (self._rand(n, A) < 0.01) & errmask.unsqueeze(0)
        ad = torch.where(sub, (ad + self._randint(1, 4, (n, A)).to(torch.uint8)) % 4, ad)
        # 0.1 % of reads: one indel inside the (inner) adapter
        m_ad = len(rc.adapter)
        r_ind = self._rand(n)
        ipos = self._randint(1, m_ad - 1, (n,))
        k = torch.arange(A, device=dev).unsqueeze(0)
        shift = torch.where((r_ind < 0.0005).unsqueeze(1) & (k >= ipos.unsqueeze(1)), 1,
                            torch.where(((r_ind >= 0.0005) & (r_ind < 0.001)).unsqueeze(1) & (k > ipos.unsqueeze(1)), -1, 0))
        ad = torch.gather(ad, 1, (k + shift).clamp(0, A - 1))
        tail = self._randint(0, 4, (n, L)).to(torch.uint8)
        j = torch.arange(L, device=dev).unsqueeze(0)
        il = ilen.unsqueeze(1)
        comb = torch.cat([ins, ad, tail], dim=1)
        gi = torch.where(j < il, j, torch.where(j < il + A, IM + (j - il), IM + A + j))
        seq = torch.gather(comb, 1, gi)
        q0 = self._randint(30, 41, (n, 1)).to(torch.float32)
        q1 = self._randint(2, 26, (n, 1)).to(torch.float32)
        q = q0 + (q1 - q0) * (j.to(torch.float32) / max(L - 1, 1)) + self._randint(-5, 6, (n, L)).to(torch.float32)
        q = q.round().clamp_(2, 41).to(torch.uint8)
        if rc.polyg_frac > 0:
            pg = (self._rand(n) < rc.polyg_frac).unsqueeze(1) & (j >= il + m_ad)
            seq = torch.where(pg, torch.full_like(seq, 2), seq)
            q = torch.where(pg, self._randint(30, 41, (n, L)).to(torch.uint8), q)
        lut = torch.tensor(list(b"ACGTN"), dtype=torch.uint8, device=dev)
        seq = lut[seq.to(torch.int64)]
        # 0.2 % of reads get one N
        hasn = self._rand(n) < 0.002
        npos = self._randint(0, L, (n,))
        seq[hasn, npos[hasn]] = ord("N")
        q = q + 33
        return self._assemble(seq, q, first_index)

    def _assemble(self, seq: torch.Tensor, q: torch.Tensor, first_index: int) -> torch.Tensor:
        dev = self.device
        n, L = seq.shape
        idx = torch.arange(first_index, first_index + n, device=dev, dtype=torch.int64)
        pre = torch.tensor(list(("@SYN%d." % self.rc.cfg_id).encode()), dtype=torch.uint8, device=dev)
        post = torch.tensor(list((" length=%d\n" % L).encode()), dtype=torch.uint8, device=dev)
        plus = torch.tensor(list(b"\n+\n"), dtype=torch.uint8, device=dev)
        nl = torch.tensor([10], dtype=torch.uint8, device=dev)
        parts = []
        lo = first_index
        end = first_index + n
        while lo < end:
            nd = len(str(lo))
            hi = min(end, 10 ** nd)
            sl = slice(lo - first_index, hi - first_index)
            k = hi - lo
            pw = 10 ** torch.arange(nd - 1, -1, -1, device=dev, dtype=torch.int64)
            digits = ((idx[sl].unsqueeze(1) // pw) % 10 + 48).to(torch.uint8)
            sp = torch.full((k, 1), 32, dtype=torch.uint8, device=dev)
            rec = torch.cat([pre.expand(k, -1), digits, sp, digits, post.expand(k, -1), seq[sl], plus.expand(k, -1), q[sl],
                             nl.expand(k, -1)], dim=1)
            parts.append(rec.reshape(-1))
            lo = hi
        return torch.cat(parts)

    def fastq(self, n_reads: int, block: int = 1 << 21, first_index: int = 0) -> torch.Tensor:
        out = []
        done = 0
        while done < n_reads:
            k = min(block, n_reads - done)
            out.append(self.block(k, first_index + done))
            done += k
        return torch.cat(out) if len(out) != 1 else out[0]


def trim_config_for(cfg_id: int, count_mode: str = "head"):
    """miRge options each benchmark config is run with."""
    from . import params as P

    if cfg_id == 2:
        return P.TrimConfig(adapters=[("back", ILLUMINA)], nextseq_trim=20, quality_cutoff="20", count_mode=count_mode)
    if cfg_id == 3:
        return P.TrimConfig(adapters=[("back", QIA_INNER)], uniq_mol_ids="0,12", qiagenumi=True, count_mode=count_mode)
    return P.TrimConfig(adapters=[("back", ILLUMINA)], count_mode=count_mode)
