"""Host-side equivalent of ``stipulate(args)`` (mirge/libs/digest.py:59-101) and of the worker
globals ``baking`` publishes (digest.py:110-122): resolve miRge's ``args`` namespace into the
plain ``mirge_trim_params`` structure the kernels consume."""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from . import abi

IUPAC = {
    "A": 1, "C": 2, "G": 4, "T": 8, "U": 8,
    "R": 5, "Y": 10, "S": 6, "W": 9, "K": 12, "M": 3,
    "B": 14, "D": 13, "H": 11, "V": 7, "N": 15, "X": 0,
}

# __main__.py:65-83: the "illumina" alias is expanded before baking() is called; accepted here too
# so that the entry points can be driven directly with a parseArg()-style namespace.
ILLUMINA_BACK = "TGGAATTCTCGGGTGCCAAGGAACTCCAG"
ILLUMINA_FRONT = "GTTCAGAGTTCTACAGTCCGACGATC"


class UnsupportedAdapterSpec(RuntimeError):
    pass


def parse_cutoffs(s) -> List[int]:
    """digest.py:19-35 (``exit`` replaced by an exception)."""
    try:
        cutoffs = [int(value) for value in str(s).split(",")]
    except ValueError as e:
        raise RuntimeError("Quality cutoff value not recognized: {}".format(e))
    if len(cutoffs) == 1:
        cutoffs = [0, cutoffs[0]]
    elif len(cutoffs) != 2:
        raise RuntimeError("Expected one value or two values separated by comma for the quality cutoff")
    return cutoffs


@dataclass
class AdapterSpec:
    where: str  # "back" | "front" | "linked" (sequence = the 5' half, sequence2 = the 3' half)
    sequence: str
    sequence2: str = ""


def parse_adapter_spec(kind: str, spec: str) -> AdapterSpec:
    """The subset of cutadapt's adapter specification language reachable from miRge's ``-a``/``-g``
    (parse.py:74-77) that this path implements: a plain (non-anchored) 3' or 5' adapter, optionally ``name=SEQ``,
    and the linked form the reference documents, ``-g "ADAPTER5...ADAPTER3"`` (docs/source/quick_start.md:208-220;
    cutadapt: both halves non-anchored, both required).  ``-a "A...B"`` anchors its 5' half and is, like everything else,
    rejected instead of silently diverging."""
    if kind not in ("back", "front"):
        raise UnsupportedAdapterSpec("adapter type %r is not supported" % (kind,))
    s = spec.strip()
    if "..." in s and kind == "front":
        body = s.split("=", 1)[1] if "=" in s else s
        halves = body.split("...")
        if len(halves) != 2 or not halves[0] or not halves[1]:
            raise UnsupportedAdapterSpec("linked adapter specification %r: expected ADAPTER5...ADAPTER3" % spec)
        if any(x.lower() == "illumina" for x in halves):
            # quick_start.md:219-221: the alias is not decoded inside a linked specification (cutadapt would take the
            # letters as bases)
            raise UnsupportedAdapterSpec("linked adapter specification %r: give the complete adapter sequences" % spec)
        five, three = (parse_adapter_spec("front", halves[0]), parse_adapter_spec("back", halves[1]))
        return AdapterSpec("linked", five.sequence, three.sequence)
    if s.lower() == "illumina":
        s = ILLUMINA_BACK if kind == "back" else ILLUMINA_FRONT
    if "=" in s:
        s = s.split("=", 1)[1]
    if s.startswith("file:") or "..." in s or ";" in s or s.startswith("^") or s.endswith("$"):
        raise UnsupportedAdapterSpec(
            "adapter specification %r (linked / anchored / file: / ;parameters) is not supported "
            "by the B200 path" % spec
        )
    s = s.upper().replace("U", "T")
    if s.endswith("X") or s.startswith("X"):
        raise UnsupportedAdapterSpec("non-internal adapter specification %r is not supported" % spec)
    if not s:
        raise UnsupportedAdapterSpec("empty adapter")
    if len(s) > abi.MAX_ADAPTER_LEN:
        raise UnsupportedAdapterSpec("adapter longer than %d nt" % abi.MAX_ADAPTER_LEN)
    bad = set(s) - set(IUPAC)
    if bad:
        raise UnsupportedAdapterSpec("adapter %r has non-IUPAC characters %s" % (spec, sorted(bad)))
    return AdapterSpec(kind, s)


@dataclass
class TrimConfig:
    """Python-level view of the trim parameters (defaults = parse.py defaults)."""

    adapters: Sequence[Tuple[str, str]] = ()
    error_rate: float = 0.12
    overlap: int = 3
    indels: bool = True
    match_adapter_wildcards: bool = True
    match_read_wildcards: bool = False
    times: int = 1
    action: str = "trim"
    nextseq_trim: Optional[int] = None
    quality_cutoff: Optional[str] = "10"
    quality_base: int = 33
    trim_n: bool = False
    cut: Sequence[int] = ()
    minimum_length: int = 16
    uniq_mol_ids: Optional[str] = None
    qiagenumi: bool = False
    count_mode: str = "head"
    cutadapt_compat: str = ""  # "2-3" | "4" ("" = MIRGE_B200_CUTADAPT_COMPAT, default "2-3"); see include/mirge_b200.h

    @classmethod
    def from_args(cls, args, count_mode: str = "head") -> "TrimConfig":
        """Read the same attributes ``stipulate``/``baking`` read from ``args``."""
        g = lambda n, d: getattr(args, n, d)
        return cls(
            adapters=list(g("adapters", [])),
            error_rate=float(g("error_rate", 0.12)),
            overlap=int(g("overlap", 3)),
            indels=bool(g("indels", True)),
            match_adapter_wildcards=bool(g("match_adapter_wildcards", True)),
            match_read_wildcards=bool(g("match_read_wildcards", False)),
            times=int(g("times", 1)),
            action=g("action", "trim"),
            nextseq_trim=g("nextseq_trim", None),
            quality_cutoff=g("quality_cutoff", "10"),
            quality_base=int(g("phred64", 33)),
            trim_n=bool(g("trim_n", False)),
            cut=list(g("cut", []) or []),
            minimum_length=int(g("minimum_length", 16)),
            uniq_mol_ids=g("uniq_mol_ids", None),
            qiagenumi=bool(g("qiagenumi", False)),
            count_mode=count_mode,
        )

    def umi(self) -> Optional[Tuple[int, int]]:
        if not self.uniq_mol_ids:
            return None
        parts = str(self.uniq_mol_ids).split(",")
        return int(parts[0]), int(parts[1])


def build_adapter(spec: AdapterSpec, cfg: TrimConfig) -> abi.Adapter:
    a = abi.Adapter()
    seq = spec.sequence
    m = len(seq)
    wildcard_ref = cfg.match_adapter_wildcards and not set(seq) <= set("ACGT")
    if not wildcard_ref and not set(seq) <= set("ACGT"):
        raise UnsupportedAdapterSpec("IUPAC adapter characters with -N (no adapter wildcards) are not supported")
    a.where = 0 if spec.where == "back" else 1
    a.link = 0
    a.m = m
    a.min_overlap = min(int(cfg.overlap), m)  # cutadapt adapters.py: min_overlap = min(min_overlap, len(sequence))
    a.indel_cost = 1 if cfg.indels else 100000
    a.wildcard_ref = 1 if wildcard_ref else 0
    a.k = int(cfg.error_rate * m)
    c = 0
    for i, ch in enumerate(seq):
        a.n_counts[i] = c
        if ch == "N":
            c += 1
        a.mask[i] = IUPAC[ch]
        a.ascii[i] = ord(ch)
    a.n_counts[m] = c
    a.effective_length = m - c if wildcard_ref else m
    if a.effective_length == 0:
        raise UnsupportedAdapterSpec("Cannot have only N wildcards in the sequence")
    for L in range(abi.MAX_ADAPTER_LEN + 1):
        a.max_err[L] = int(L * cfg.error_rate)  # floor of the double product, as cutadapt compares
    return a


def build_trim_params(cfg: TrimConfig) -> abi.TrimParams:
    """``stipulate`` + the worker globals, flattened (modifier order: digest.py:87-99)."""
    if cfg.action != "trim":
        raise RuntimeError("action=%r is not supported (miRge always uses 'trim', parse.py:95)" % cfg.action)
    if cfg.match_read_wildcards:
        raise RuntimeError("--match-read-wildcards is not supported")
    p = abi.TrimParams()
    specs = []
    links = []  # per flattened adapter: mirge_adapter.link
    for (k, s) in cfg.adapters:
        sp = parse_adapter_spec(k, s)
        if sp.where == "linked":  # two consecutive entries: the 5' half points at the 3' half
            specs += [AdapterSpec("front", sp.sequence), AdapterSpec("back", sp.sequence2)]
            links += [len(specs), abi.LINK_BACK_HALF]  # (1 + index of the 3' half) = len(specs) after both were appended
        else:
            specs.append(sp)
            links.append(0)
    if any(links) and cfg.qiagenumi:
        raise UnsupportedAdapterSpec("linked adapters together with --qiagenumi are not supported")
    if len(specs) > abi.MAX_ADAPTERS:
        raise RuntimeError("at most %d adapters are supported" % abi.MAX_ADAPTERS)
    mods = []
    if cfg.nextseq_trim is not None:
        mods.append((abi.MOD_NEXTSEQ, int(cfg.nextseq_trim), cfg.quality_base, 0))
    if cfg.quality_cutoff is not None:
        q5, q3 = parse_cutoffs(cfg.quality_cutoff)
        mods.append((abi.MOD_QUALITY, q5, q3, cfg.quality_base))
    if specs:
        mods.append((abi.MOD_ADAPTER, 0, 0, 0))
    if cfg.trim_n:
        mods.append((abi.MOD_NEND, 0, 0, 0))
    cut = list(cfg.cut or [])
    if cut:
        if len(cut) > 2:
            raise RuntimeError("You cannot remove bases from more than two ends.")  # digest.py:47-48
        if len(cut) == 2 and cut[0] * cut[1] > 0:
            raise RuntimeError("You cannot remove bases from the same end twice.")  # digest.py:49-50
        for c in cut:
            if c != 0:
                mods.append((abi.MOD_CUT, int(c), 0, 0))
    if len(mods) > abi.MAX_MODS:
        raise RuntimeError("too many modifiers")
    p.n_mods = len(mods)
    for i, (k, a, b, c) in enumerate(mods):
        p.mod_kind[i], p.mod_a[i], p.mod_b[i], p.mod_c[i] = k, a, b, c
    p.n_adapters = len(specs)
    for i, s in enumerate(specs):
        p.adapters[i] = build_adapter(s, cfg)
        p.adapters[i].link = links[i]
    p.times = int(cfg.times)
    p.min_len = int(cfg.minimum_length)
    umi = cfg.umi()
    if cfg.qiagenumi:
        if umi is None:
            raise RuntimeError("--qiagenumi requires -umi x,y")
        if not specs:
            raise RuntimeError("--qiagenumi requires the internal adapter (-a)")
        p.umi_mode = abi.UMI_QIAGEN
        p.qia_adapter_len = len(str(cfg.adapters[0][1])) if str(cfg.adapters[0][1]).lower() != "illumina" else specs[0].sequence.__len__()
    elif umi is not None:
        p.umi_mode = abi.UMI_FLANKS
    else:
        p.umi_mode = abi.UMI_NONE
    if umi is not None:
        p.umi5, p.umi3 = umi
    if cfg.count_mode not in ("head", "release"):
        raise RuntimeError("count_mode must be 'head' or 'release'")
    p.count_mode = abi.COUNT_HEAD if cfg.count_mode == "head" else abi.COUNT_RELEASE
    compat = cfg.cutadapt_compat or os.environ.get("MIRGE_B200_CUTADAPT_COMPAT", "2-3")
    if compat not in ("2-3", "4"):
        raise RuntimeError("cutadapt_compat must be '2-3' or '4'")
    p.compat = abi.COMPAT_CUTADAPT4 if compat == "4" else abi.COMPAT_CUTADAPT23
    return p


def trim_slots(p: abi.TrimParams) -> int:
    """Emission slots per read (mirrors mirge_trim_slots)."""
    return p.n_mods if (p.umi_mode != abi.UMI_QIAGEN and p.count_mode == abi.COUNT_HEAD) else 1
