"""Host-side equivalent of ``stipulate(args)`` (mirge/libs/digest.py:59-101) and of the worker
globals ``baking`` publishes (digest.py:110-122): resolve miRge's ``args`` namespace into the
plain ``mirge_trim_params`` structure the kernels consume."""
from __future__ import annotations

import os

from dataclasses import dataclass, field
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
    """One parsed ``-a`` / ``-g`` specification.  ``where`` is a key of ``abi.WHERE`` (cutadapt's ``Where`` of the placement)
    or "linked" (sequence / where5 / params = the 5' half, sequence2 / where2 / params2 = the 3' half)."""

    where: str
    sequence: str
    sequence2: str = ""
    params: dict = field(default_factory=dict)  # per-adapter search parameters: max_error_rate, min_overlap, indels
    where5: str = "front"  # linked pairs: placement of the halves
    where2: str = "back"
    params2: dict = field(default_factory=dict)
    front_required: bool = True
    back_required: bool = True


# cutadapt parser.py AdapterSpecification.allowed_parameters (abbreviation -> parameter); "indels" / "noindels" are the
# per-adapter switches later releases added
_PARAMETERS = {"e": "max_error_rate", "error_rate": "max_error_rate", "max_errors": "max_error_rate", "o": "min_overlap",
               "max_error_rate": None, "min_overlap": None, "anywhere": None, "required": None, "optional": None,
               "indels": None, "noindels": None}


def expand_braces(seq: str) -> str:
    """cutadapt ``AdapterSpecification.expand_braces``: ``x{n}`` stands for n times the character x."""
    out = []
    i = 0
    while i < len(seq):
        c = seq[i]
        if c == "}":
            raise UnsupportedAdapterSpec('adapter %r: "}" cannot be used here' % seq)
        if c == "{":
            j = seq.find("}", i)
            if not out or j < 0 or not seq[i + 1 : j].isdigit() or not 0 <= int(seq[i + 1 : j]) <= 10000:
                raise UnsupportedAdapterSpec('adapter %r: "{n}" must follow a character and hold a number' % seq)
            last = out.pop()
            out.append(last * int(seq[i + 1 : j]))
            i = j + 1
            continue
        out.append(c)
        i += 1
    return "".join(out)


def _parse_parameters(text: str, spec: str) -> dict:
    """``key=value;key;...`` behind the first ';' of a specification (cutadapt ``_parse_parameters``)."""
    out = {}
    for fld in text.split(";"):
        fld = fld.strip()
        if not fld:
            continue
        key, eq, value = fld.partition("=")
        key, value = key.strip(), value.strip()
        if eq and not value:
            raise UnsupportedAdapterSpec("adapter specification %r: no value given for %r" % (spec, key))
        if key not in _PARAMETERS:
            raise UnsupportedAdapterSpec("adapter specification %r: unknown parameter %r" % (spec, key))
        key = _PARAMETERS[key] or key
        if not eq:
            val = True
        else:
            try:
                val = int(value)
            except ValueError:
                try:
                    val = float(value)
                except ValueError:
                    raise UnsupportedAdapterSpec("adapter specification %r: %r is not a number" % (spec, value))
        if key in out:
            raise UnsupportedAdapterSpec("adapter specification %r: parameter %r given twice" % (spec, key))
        out[key] = val
    if "optional" in out and "required" in out:
        raise UnsupportedAdapterSpec("adapter specification %r: 'optional' and 'required' cannot be specified at the same time" % spec)
    if "indels" in out and "noindels" in out:
        raise UnsupportedAdapterSpec("adapter specification %r: 'indels' and 'noindels' cannot be specified at the same time" % spec)
    if "optional" in out:
        out["required"] = False
        del out["optional"]
    if "noindels" in out:
        out["indels"] = False
        del out["noindels"]
    if out.get("anywhere"):
        raise UnsupportedAdapterSpec("adapter specification %r: ;anywhere (an adapter on either end) is not supported" % spec)
    out.pop("anywhere", None)
    e = out.get("max_error_rate")
    if e is not None and not 0 <= e < 1:
        raise UnsupportedAdapterSpec("adapter specification %r: the error rate must be in [0, 1)" % spec)
    if "min_overlap" in out and (out["min_overlap"] is True or int(out["min_overlap"]) < 1):
        raise UnsupportedAdapterSpec("adapter specification %r: min_overlap must be at least 1" % spec)
    return out


def _parse_half(spec: str, kind: str):
    """cutadapt ``AdapterSpecification.parse``: ``[name=][^|X]SEQ[$|X][;parameters]`` -> (where, sequence, parameters)."""
    body, _, ptext = spec.partition(";")
    if "=" in body:
        body = body.split("=", 1)[1]
    body = body.strip()
    params = _parse_parameters(ptext, spec)
    if body.lower() == "illumina":  # __main__.py:65-83 expands the alias before baking() is called; accepted here too
        body = ILLUMINA_BACK if kind == "back" else ILLUMINA_FRONT
    body = expand_braces(body)
    if body and not body.strip("Xx"):
        raise UnsupportedAdapterSpec("adapter specification %r consists of X only" % spec)
    front = back = None
    if body.startswith("^"):
        front, body = "anchored", body[1:]
    if body.upper().startswith("X"):
        if front:
            raise UnsupportedAdapterSpec('adapter specification %r: either "^" or "X" must be used, not both' % spec)
        front, body = "noninternal", body.lstrip("xX")
    if body.endswith("$"):
        back, body = "anchored", body[:-1]
    if body.upper().endswith("X"):
        if back:
            raise UnsupportedAdapterSpec('adapter specification %r: either "$" or "X" must be used, not both' % spec)
        back, body = "noninternal", body.rstrip("xX")
    if front and back:
        raise UnsupportedAdapterSpec("adapter specification %r: only one placement restriction is possible" % spec)
    if kind == "front" and back:
        raise UnsupportedAdapterSpec("adapter specification %r: a 5' adapter takes XADAPTER or ^ADAPTER" % spec)
    if kind == "back" and front:
        raise UnsupportedAdapterSpec("adapter specification %r: a 3' adapter takes ADAPTERX or ADAPTER$" % spec)
    restriction = front or back
    where = {None: kind, "anchored": "prefix" if kind == "front" else "suffix",
             "noninternal": kind + "_not_internal"}[restriction]
    seq = body.upper().replace("U", "T")
    if not seq:
        raise UnsupportedAdapterSpec("empty adapter")
    if len(seq) > abi.MAX_ADAPTER_LEN:
        raise UnsupportedAdapterSpec("adapter longer than %d nt" % abi.MAX_ADAPTER_LEN)
    bad = set(seq) - set(IUPAC)
    if bad:
        raise UnsupportedAdapterSpec("adapter %r has non-IUPAC characters %s" % (spec, sorted(bad)))
    return where, seq, params, restriction


def parse_adapter_spec(kind: str, spec: str) -> AdapterSpec:
    """One specification of cutadapt's adapter language as it reaches ``stipulate`` through miRge's ``-a`` / ``-g``
    (parse.py:74-77; cutadapt parser.py ``AdapterParser._parse``): plain, anchored (``^SEQ``, ``SEQ$``) and non-internal
    (``XSEQ``, ``SEQX``) adapters, ``name=``, ``x{n}`` repeats, per-adapter ``;parameters`` and the linked form
    ``ADAPTER5...ADAPTER3`` the reference documents (docs/source/quick_start.md:208-220) -- with ``-g`` both halves are
    required, with ``-a`` a half is required only when it is anchored, ``;required`` / ``;optional`` override both.
    ``file:`` specifications hold several adapters: ``parse_adapter_specs``.  What the kernels have no form for
    (``;anywhere``, adapters on either end) raises instead of silently diverging."""
    if kind not in ("back", "front"):
        raise UnsupportedAdapterSpec("adapter type %r is not supported" % (kind,))
    s = spec.strip()
    if s.startswith("file:"):
        raise UnsupportedAdapterSpec("adapter specification %r names a file of adapters: use parse_adapter_specs" % spec)
    one, dots, two = s.partition("...")
    if dots and one and two:
        if "..." in two:
            raise UnsupportedAdapterSpec("linked adapter specification %r: expected ADAPTER5...ADAPTER3" % spec)
        if any(x.partition(";")[0].split("=")[-1].strip().lower() == "illumina" for x in (one, two)):
            # quick_start.md:219-221: the alias is not decoded inside a linked specification (cutadapt would take the
            # letters as bases)
            raise UnsupportedAdapterSpec("linked adapter specification %r: give the complete adapter sequences" % spec)
        w5, s5, p5, r5 = _parse_half(one, "front")
        w3, s3, p3, r3 = _parse_half(two, "back")
        # cutadapt _parse_linked: -g needs both halves, -a only the anchored ones; ;required / ;optional decide otherwise
        req5 = True if kind == "front" else r5 == "anchored"
        req3 = True if kind == "front" else r3 == "anchored"
        req5 = bool(p5.pop("required", req5))
        req3 = bool(p3.pop("required", req3))
        return AdapterSpec("linked", s5, s3, p5, w5, w3, p3, req5, req3)
    if dots:
        if not one and kind == "back":  # -a ...ADAPTER: a plain 3' adapter
            s = two
        elif not two:  # -a ADAPTER... / -g ADAPTER...: a plain 5' adapter
            s, kind = one, "front"
        else:
            raise UnsupportedAdapterSpec("invalid adapter specification %r" % spec)
    where, seq, params, _ = _parse_half(s, kind)
    if "required" in params:
        raise UnsupportedAdapterSpec("adapter specification %r: 'optional' and 'required' belong to linked adapters" % spec)
    return AdapterSpec(where, seq, params=params)


def read_adapter_fasta(path: str) -> List[str]:
    """The sequences of a FASTA file of adapters (``-a file:adapters.fa``; cutadapt reads it with dnaio's FastaReader:
    '>' header lines, sequences possibly over several lines, '#' comment lines and blank lines skipped)."""
    seqs: List[str] = []
    cur: Optional[List[str]] = None
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if not line or (line.startswith("#") and cur is None):
                continue
            if line.startswith(">"):
                if cur is not None:
                    seqs.append("".join(cur))
                cur = []
            elif cur is None:
                raise UnsupportedAdapterSpec("%s: not a FASTA file (no '>' before the first sequence)" % path)
            else:
                cur.append(line)
    if cur is not None:
        seqs.append("".join(cur))
    return seqs


def parse_adapter_specs(kind: str, spec: str) -> List[AdapterSpec]:
    """``parse_adapter_spec`` for every adapter a ``-a`` / ``-g`` argument stands for: one, or with ``file:PATH`` one per
    FASTA record (cutadapt parser.py ``AdapterParser.parse``: each record's sequence is a specification of its own)."""
    s = spec.strip()
    if s.startswith("file:"):
        try:
            seqs = read_adapter_fasta(s[5:])
        except OSError as e:
            raise UnsupportedAdapterSpec("adapter specification %r: %s" % (spec, e))
        if not seqs:
            raise UnsupportedAdapterSpec("adapter specification %r: the file holds no adapter" % spec)
        return [parse_adapter_spec(kind, q) for q in seqs]
    return [parse_adapter_spec(kind, s)]


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


def build_adapter(spec: AdapterSpec, cfg: TrimConfig, half: int = 0) -> abi.Adapter:
    """One plain adapter, or with ``half`` 1 / 2 the 5' / 3' half of a linked pair, in the Aligner terms the kernels take."""
    a = abi.Adapter()
    seq, where, prm = ((spec.sequence, spec.where, spec.params) if half == 0 else
                       (spec.sequence, spec.where5, spec.params) if half == 1 else (spec.sequence2, spec.where2, spec.params2))
    m = len(seq)
    rate = float(prm.get("max_error_rate", cfg.error_rate))
    overlap = int(prm.get("min_overlap", cfg.overlap))
    indels = bool(prm.get("indels", cfg.indels))
    wildcard_ref = cfg.match_adapter_wildcards and not set(seq) <= set("ACGT")
    if not wildcard_ref and not set(seq) <= set("ACGT"):
        raise UnsupportedAdapterSpec("IUPAC adapter characters with -N (no adapter wildcards) are not supported")
    a.where = abi.WHERE[where]
    a.link = 0
    a.m = m
    a.min_overlap = min(overlap, m)  # cutadapt adapters.py: min_overlap = min(min_overlap, len(sequence))
    if where in ("prefix", "suffix"):
        a.min_overlap = m  # anchored adapters occur in full (every alignment the placement admits spans the adapter)
    a.indel_cost = 1 if indels else 100000
    a.wildcard_ref = 1 if wildcard_ref else 0
    a.wildcard_read = 1 if cfg.match_read_wildcards else 0
    a.k = int(rate * m)
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
        a.max_err[L] = int(L * rate)  # floor of the double product, as cutadapt compares
    return a


def flatten_adapters(cfg: TrimConfig):
    """``args.adapters`` as the flat list the kernels walk: [(AdapterSpec, half)] plus ``mirge_adapter.link`` per entry.
    A linked pair takes two consecutive entries, its 5' half pointing at its 3' half."""
    flat, links = [], []
    for (k, s) in cfg.adapters:
        for sp in parse_adapter_specs(k, s):
            if sp.where == "linked":
                flat += [(sp, 1), (sp, 2)]
                link = len(flat)  # 1 + index of the 3' half
                if not sp.front_required:
                    link |= abi.LINK_FRONT_OPTIONAL
                if not sp.back_required:
                    link |= abi.LINK_BACK_OPTIONAL
                links += [link, abi.LINK_BACK_HALF]
            else:
                flat.append((sp, 0))
                links.append(0)
    return flat, links


def build_trim_params(cfg: TrimConfig) -> abi.TrimParams:
    """``stipulate`` + the worker globals, flattened (modifier order: digest.py:87-99)."""
    if cfg.action != "trim":
        raise RuntimeError("action=%r is not supported (miRge always uses 'trim', parse.py:95)" % cfg.action)
    p = abi.TrimParams()
    specs, links = flatten_adapters(cfg)
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
    for i, (sp, half) in enumerate(specs):
        p.adapters[i] = build_adapter(sp, cfg, half)
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
        p.qia_adapter_len = len(str(cfg.adapters[0][1])) if str(cfg.adapters[0][1]).lower() != "illumina" else len(specs[0][0].sequence)
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
