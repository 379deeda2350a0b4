"""CPU oracle (pure Python) for miRge3.0's digest -> collapse -> annotate hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mirge3.0_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker.

PARITY UNPINNED.  The arithmetic of this path lives in two third-party tools that are not
vendored in /root/reference and are not installable here (no network): ``cutadapt``
(unpinned in reference ``setup.py:17``; the reference says "ADOPTED FROM CUTADAPT 2.7",
``mirge/libs/digest.py:21,41,61``) with ``dnaio``/``xopen``, and the ``bowtie`` 1.x binary
(``mirge/libs/miRgeEssential.py:17``).  The reference has no tests and no golden files.  This
module restates the *published* algorithms of those tools (cutadapt 2.x-3.x ``_align.pyx``
``Aligner.locate``, ``qualtrim.pyx``, ``modifiers.py``; ``dnaio`` chunking/FASTQ parsing;
bowtie 1 manual ``-v``/``-n``/``--best --strata`` semantics) and anchors on the reference's own
call sites plus the ten known-answer reads in ``docs/source/quick_start.md:285-315``
(see ``tests/test_oracle_golden.py``).

Every function cites the reference ``file:line`` it follows (paths relative to /root/reference).
Pure-Python loops: use for small cases only; ``oracle/mirge_oracle.c`` is the fast restatement,
validated against this file.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# FASTQ chunking + parsing  (dnaio; called from mirge/libs/digest.py:136,140,324)
# --------------------------------------------------------------------------------------


class FastqFormatError(ValueError):
    """Mirrors dnaio.FastqFormatError for malformed records (digest.py:324 lets it propagate)."""


def _fastq_head(buf: bytes, end: int) -> int:
    """dnaio.chunks._fastq_head: offset just after the last complete 4-line record."""
    linebreaks = buf.count(b"\n", 0, end)
    right = end
    for _ in range(linebreaks % 4 + 1):
        right = buf.rfind(b"\n", 0, right)
    return right + 1


def read_chunks(data: bytes, buffer_size: int = 4_000_000) -> List[Tuple[int, int]]:
    """(start, end) byte ranges dnaio.read_chunks(f, buffer_size) yields for a *plain* file
    (digest.py:140; args.buffer_size default parse.py:100).  For a regular uncompressed file
    every ``readinto`` fills the buffer, so the boundaries are a pure function of the bytes:
    each chunk is the longest prefix of the next ``buffer_size`` bytes that ends on a record
    boundary; the final partial tail is yielded as-is at EOF."""
    if not data:
        return []
    if data[0:1] != b"@":
        raise FastqFormatError("Input file format unknown (first byte is not '@')")
    out = []
    s = 0
    n = len(data)
    while s < n:
        window_end = min(n, s + buffer_size)
        if window_end == n:
            # last fill: records that complete inside it are yielded, the tail afterwards
            e = s + _fastq_head(data[s:window_end], window_end - s)
            if e > s:
                out.append((s, e))
            if e < n:
                if e == s and n - s >= buffer_size:
                    raise OverflowError("FASTA/FASTQ record does not fit into buffer")
                out.append((e, n))
            break
        e = s + _fastq_head(data[s:window_end], window_end - s)
        if e == s:
            raise OverflowError("FASTA/FASTQ record does not fit into buffer")
        out.append((s, e))
        s = e
    return out


def parse_fastq(data: bytes) -> List[Tuple[str, str, str]]:
    """dnaio FastqIter semantics (digest.py:324-325): 4-line records, trailing '\\r' stripped,
    line 1 starts with '@', line 3 with '+', len(seq) == len(qual).  Last line may lack '\\n'."""
    if not data:
        return []
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if len(lines) % 4 != 0:
        raise FastqFormatError("Premature end of file (line count %d not a multiple of 4)" % len(lines))
    recs = []
    for i in range(0, len(lines), 4):
        h, s, p, q = (x[:-1] if x.endswith(b"\r") else x for x in lines[i : i + 4])
        if not h.startswith(b"@"):
            raise FastqFormatError("Line %d expected to start with '@'" % (i + 1))
        if not p.startswith(b"+"):
            raise FastqFormatError("Line %d expected to start with '+'" % (i + 3))
        if len(s) != len(q):
            raise FastqFormatError("Length of sequence and qualities differ (record %d)" % (i // 4))
        recs.append((h[1:].decode("latin-1"), s.decode("latin-1"), q.decode("latin-1")))
    return recs


# --------------------------------------------------------------------------------------
# Quality trimming  (cutadapt qualtrim.pyx; constructed at digest.py:87-91)
# --------------------------------------------------------------------------------------


def nextseq_trim_index(seq: str, qual: str, cutoff: int, base: int = 33) -> int:
    """cutadapt.qualtrim.nextseq_trim_index (NextseqQualityTrimmer, digest.py:87-88)."""
    s = 0
    max_qual = 0
    max_i = len(qual)
    for i in range(len(qual) - 1, -1, -1):
        q = ord(qual[i]) - base
        if seq[i] == "G":
            q = cutoff - 1
        s += cutoff - q
        if s < 0:
            break
        if s > max_qual:
            max_qual = s
            max_i = i
    return max_i


def quality_trim_index(qual: str, cutoff_front: int, cutoff_back: int, base: int = 33) -> Tuple[int, int]:
    """cutadapt.qualtrim.quality_trim_index (QualityTrimmer, digest.py:89-91)."""
    s = 0
    max_qual = 0
    start = 0
    stop = len(qual)
    for i in range(len(qual)):
        s += cutoff_front - (ord(qual[i]) - base)
        if s < 0:
            break
        if s > max_qual:
            max_qual = s
            start = i + 1
    max_qual = 0
    s = 0
    for i in range(len(qual) - 1, -1, -1):
        s += cutoff_back - (ord(qual[i]) - base)
        if s < 0:
            break
        if s > max_qual:
            max_qual = s
            stop = i
    if start >= stop:
        start, stop = 0, 0
    return start, stop


# --------------------------------------------------------------------------------------
# Adapter alignment  (cutadapt _align.pyx Aligner.locate; AdapterCutter at digest.py:93-96)
# --------------------------------------------------------------------------------------

IUPAC = {
    "A": 1, "C": 2, "G": 4, "T": 8, "U": 8,
    "R": 1 | 4, "Y": 2 | 8, "S": 2 | 4, "W": 1 | 8, "K": 4 | 8, "M": 1 | 2,
    "B": 2 | 4 | 8, "D": 1 | 4 | 8, "H": 1 | 2 | 8, "V": 1 | 2 | 4, "N": 15, "X": 0,
}
ACGT = {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8}
ACGT_ASCII = {"A": 1, "C": 2, "G": 4, "T": 8}  # plain character comparison: a U in the read is not a T

INDEL_OFF_COST = 100000  # cutadapt adapters.py: "indel_cost = 1 if self.indels else 100000"

# cutadapt adapters.py ``Where``: which ends of the alignment are free -- (start within the adapter, stop within the
# adapter, start within the read, stop within the read); "within the adapter" = the adapter's start / end may be skipped
WHERE_FLAGS = {
    "back": (False, True, True, True),  # -a SEQ: anywhere in the read, the adapter may hang over the 3' end
    "front": (True, False, True, True),  # -g SEQ: anywhere, may hang over the 5' end
    "prefix": (False, False, False, True),  # -g ^SEQ (anchored 5'): read and adapter start together
    "suffix": (False, False, True, False),  # -a SEQ$ (anchored 3'): read and adapter end together
    "front_not_internal": (True, False, False, True),  # -g XSEQ: as front, but never inside the read
    "back_not_internal": (False, True, True, False),  # -a SEQX
}
REMOVE_BEFORE = ("front", "prefix", "front_not_internal")  # 5' forms: what precedes the match goes too


@dataclass
class Adapter:
    """One parsed adapter of cutadapt's specification language as miRge's CLI hands it over (``-a SPEC`` / ``-g SPEC``,
    parse.py:74-77): ``where`` is cutadapt's ``Where`` of the placement -- plain 3' / 5' adapters, the anchored forms
    ``^SEQ`` / ``SEQ$`` and the non-internal forms ``XSEQ`` / ``SEQX``."""

    where: str  # one of WHERE_FLAGS
    sequence: str
    max_error_rate: float = 0.12  # parse.py:91
    min_overlap: int = 3  # parse.py:90
    indels: bool = True  # parse.py:101
    adapter_wildcards: bool = True  # parse.py:98 (only effective if sequence has non-ACGT)
    read_wildcards: bool = False  # parse.py:97 (--match-read-wildcards: IUPAC characters of the READ match as sets)

    def __post_init__(self):
        self.sequence = self.sequence.upper().replace("U", "T")
        if not self.sequence:
            raise ValueError("Adapter sequence is empty")
        self.wildcard_ref = self.adapter_wildcards and not set(self.sequence) <= set("ACGT")
        if not self.wildcard_ref and not set(self.sequence) <= set("ACGT"):
            raise ValueError("non-ACGT adapter characters need adapter wildcards (unsupported combination)")
        m = len(self.sequence)
        self.min_overlap = min(int(self.min_overlap), m)  # cutadapt adapters.py: an adapter shorter than -O lowers it
        self.n_counts = [0] * (m + 1)
        c = 0
        for i, ch in enumerate(self.sequence):
            self.n_counts[i] = c
            if ch == "N":
                c += 1
        self.n_counts[m] = c
        self.effective_length = m - c if self.wildcard_ref else m
        if self.effective_length == 0:
            raise ValueError("Cannot have only N wildcards in the sequence")
        self.masks = [IUPAC[ch] for ch in self.sequence]
        if self.where not in WHERE_FLAGS:
            raise ValueError("unknown adapter type %r" % (self.where,))
        if self.where in ("prefix", "suffix"):
            self.min_overlap = m  # cutadapt adapters.py: anchored adapters must occur in full


@dataclass
class LinkedAdapter:
    """cutadapt ``LinkedAdapter`` ("ADAPTER1...ADAPTER2", quick_start.md:208-220): a 5' adapter, then a 3' adapter
    searched in what the 5' match leaves.  ``-g A...B``: both required, A not anchored; ``-a A...B``: A anchored and
    required, B optional (cutadapt parser.py ``_parse_linked``); ``^`` / ``$`` anchor explicitly."""

    front: Adapter
    back: Adapter
    front_required: bool = True
    back_required: bool = True

    @property
    def where(self):
        return "linked"


def _allowed(length: int, rate: float) -> float:
    return length * rate  # evaluated in double exactly as cutadapt's "cost <= length * max_error_rate"


def locate(ad: Adapter, read: str, compat: str = "2-3") -> Optional[Tuple[int, int, int, int, int, int]]:
    """cutadapt ``Aligner.locate(query)`` for back (3') and front (5') adapters.

    ``compat`` "2-3" (the versions the reference names): every cell carries its number of matches and a candidate
    wins with more matches, then lower cost.  "4": cutadapt >= 4.0 carries a score instead -- match +1, mismatch -1,
    insertion / deletion -2, the first column of a 3' adapter starts at -2 i -- and a candidate wins with a higher
    score, then lower cost; the fifth field of the result is then the score.  (Restated from the published
    description; parity with an install is what tests/test_tier3_real_tools.py checks when one is there.)

    Returns (astart, astop, rstart, rstop, matches, errors) or None.  Unit-cost semi-global DP,
    adapter = rows, read = columns, one column kept; each cell carries (cost, origin, matches).
    Cell choice: equal characters -> diagonal; else mismatch if cd<=cdel and cd<=cins, else
    insertion if cins<=cdel, else deletion.  Candidates: row m after every column, then (after the
    last column) rows first_i..m in ascending order; a candidate replaces the best iff it has more
    matches, or equally many and a lower cost.  cutadapt's Ukkonen cut-off ("last") only skips
    cells whose cost exceeds k=int(rate*m); no accepted cell depends on them, so the full DP used
    here yields identical results.
    """
    m = len(ad.sequence)
    n = len(read)
    rate = ad.max_error_rate
    ins_cost = del_cost = 1 if ad.indels else INDEL_OFF_COST
    up = read.upper()
    # read characters -> 4-bit class: a character outside ACGT never matches, unless --match-read-wildcards
    # (parse.py:97) turns the read's IUPAC characters into sets as well (_align.pyx: query translated with IUPAC_TABLE)
    # Without any wildcards _align.pyx compares the ASCII characters themselves; its translation tables -- in which a U
    # stands for T -- are only in use when adapter or read wildcards are active.
    table = IUPAC if ad.read_wildcards else (ACGT if ad.wildcard_ref else ACGT_ASCII)
    rmask = [table.get(ch, 0) for ch in up]
    start_in_ref, stop_in_ref, start_in_query, stop_in_query = WHERE_FLAGS[ad.where]
    if compat not in ("2-3", "4"):
        raise ValueError("compat must be '2-3' or '4'")
    # what a step adds to the merit a cell carries (matches, or the score of cutadapt >= 4)
    w_mis, w_indel = (-1, -2) if compat == "4" else (0, 0)
    k = int(rate * m)
    # columns that matter (Aligner.locate): an anchored start cannot use more than m + k read bases, an anchored end
    # only the last m + k
    max_n = n if start_in_query else min(n, m + k)
    min_n = 0 if stop_in_query else max(0, n - m - k)
    cost = [0] * (m + 1)
    origin = [0] * (m + 1)
    matches = [0] * (m + 1)
    for i in range(m + 1):
        if not start_in_ref and not start_in_query:
            cost[i], origin[i] = max(i, min_n) * ins_cost, 0
        elif start_in_ref and not start_in_query:
            cost[i], origin[i] = min_n * ins_cost, min(0, min_n - i)
        elif not start_in_ref and start_in_query:
            cost[i], origin[i] = i * ins_cost, max(0, min_n - i)
        else:
            cost[i], origin[i] = min(i, min_n) * ins_cost, min_n - i
        if not start_in_ref:
            matches[i] = i * w_indel
    best_cost = m + n
    best_origin = 0
    best_matches = 0 if compat != "4" else -(1 << 30)
    best_ref_stop = m
    best_query_stop = n
    stopped_early = False

    def eff_len_row_m(length):
        if ad.wildcard_ref:
            if length < m:
                return length - (ad.n_counts[m] - ad.n_counts[m - length])
            return ad.effective_length
        return length

    for j in range(min_n + 1, max_n + 1):
        diag_c, diag_o, diag_m = cost[0], origin[0], matches[0]
        if start_in_query:
            origin[0] = j
        else:
            cost[0] = j * ins_cost
            matches[0] = j * w_indel
        rc = rmask[j - 1]
        for i in range(1, m + 1):
            if ad.masks[i - 1] & rc:
                c, o, mt = diag_c, diag_o, diag_m + 1
            else:
                cd = diag_c + 1
                cdel = cost[i] + del_cost
                cins = cost[i - 1] + ins_cost
                if cd <= cdel and cd <= cins:
                    c, o, mt = cd, diag_o, diag_m + w_mis
                elif cins <= cdel:
                    c, o, mt = cins, origin[i - 1], matches[i - 1] + w_indel
                else:
                    c, o, mt = cdel, origin[i], matches[i] + w_indel
            diag_c, diag_o, diag_m = cost[i], origin[i], matches[i]
            cost[i], origin[i], matches[i] = c, o, mt
        # row m is examined only when column[m].cost <= k ("last == m"), and only if the match may end inside the read
        if cost[m] <= k and stop_in_query:
            length = m + min(origin[m], 0)
            c = cost[m]
            mt = matches[m]
            if (
                length >= ad.min_overlap
                and c <= _allowed(eff_len_row_m(length), rate)
                and (mt > best_matches or (mt == best_matches and c < best_cost))
            ):
                best_matches, best_cost, best_origin = mt, c, origin[m]
                best_ref_stop, best_query_stop = m, j
                if c == 0 and mt == m:
                    stopped_early = True
                    break
    if not stopped_early and max_n == n:
        first_i = 0 if stop_in_ref else m
        for i in range(first_i, m + 1):
            length = i + min(origin[i], 0)
            c = cost[i]
            mt = matches[i]
            if ad.wildcard_ref:
                if length < m:
                    ref_start = -min(origin[i], 0)
                    eff = length - (ad.n_counts[i] - ad.n_counts[ref_start])
                else:
                    eff = ad.effective_length
            else:
                eff = length
            if (
                length >= ad.min_overlap
                and c <= _allowed(eff, rate)
                and (mt > best_matches or (mt == best_matches and c < best_cost))
            ):
                best_matches, best_cost, best_origin = mt, c, origin[i]
                best_ref_stop, best_query_stop = i, n
    if best_cost == m + n:
        return None
    if best_origin >= 0:
        start1, start2 = 0, best_origin
    else:
        start1, start2 = -best_origin, 0
    return (start1, best_ref_stop, start2, best_query_stop, best_matches, best_cost)


def match_to(ad: Adapter, read: str, compat: str = "2-3") -> Optional[Tuple[int, int, int, int, int, int]]:
    """cutadapt ``Adapter.match_to``: exact ``str.find`` on the upper-cased read first (only when
    the adapter has no wildcards), otherwise ``Aligner.locate``.  The fast path returns what the
    DP would (leftmost exact occurrence: the DP stops at the first column with cost 0, matches m).
    """
    if isinstance(ad, LinkedAdapter):
        return match_linked(ad, read, compat)
    up = read.upper()
    if not ad.wildcard_ref:
        m = len(ad.sequence)
        pos = -1
        if ad.where == "prefix":
            pos = 0 if up.startswith(ad.sequence) else -1
        elif ad.where == "suffix":
            pos = len(up) - m if up.endswith(ad.sequence) else -1
        elif ad.where in ("back", "front"):
            pos = up.find(ad.sequence)
        # (the non-internal forms have no shortcut in cutadapt: an occurrence inside the read is not a match for them)
        if pos >= 0:
            return (0, m, pos, pos + m, m, 0)  # (m matches; also the score of m matches)
    return locate(ad, read, compat)


def match_linked(ad: LinkedAdapter, read: str, compat: str = "2-3") -> Optional[Tuple[int, int, int, int, int, int]]:
    """cutadapt ``LinkedAdapter.match_to``: the 5' adapter on the read, the 3' adapter on what the 5' match leaves
    (``read[front.rstop:]``); a missing half ends the search when that half is required (and a pair needs at least its
    5' half).  ``LinkedMatch``: matches and errors are the sums over the halves that matched.  Returned in the shape of
    the other matches, with the two cut points in the read fields: (0, 0, first base kept, end of what is kept,
    matches, errors)."""
    f = match_to(ad.front, read, compat)
    if f is None and ad.front_required:
        return None
    rest = f[3] if f is not None else 0
    b = match_to(ad.back, read[rest:], compat)
    if b is None and (ad.back_required or f is None):
        return None
    keep_stop = rest + b[2] if b is not None else len(read)
    return (0, 0, rest, keep_stop, (f[4] if f else 0) + (b[4] if b else 0), (f[5] if f else 0) + (b[5] if b else 0))


def best_match(adapters: Sequence[Adapter], read: str, compat: str = "2-3"):
    """cutadapt ``AdapterCutter._best_match``: most matches (cutadapt >= 4: the highest score) wins, then fewer
    errors; first adapter wins remaining ties."""
    best = None
    best_ad = None
    for ad in adapters:
        mt = match_to(ad, read, compat)
        if mt is None:
            continue
        if best is None or mt[4] > best[4] or (mt[4] == best[4] and mt[5] < best[5]):
            best, best_ad = mt, ad
    return best_ad, best


# --------------------------------------------------------------------------------------
# Modifier pipeline  (stipulate(), digest.py:59-101) on (start, stop) windows of the original read
# --------------------------------------------------------------------------------------


def parse_cutoffs(s: str) -> List[int]:
    """digest.py:19-35."""
    cutoffs = [int(v) for v in str(s).split(",")]
    if len(cutoffs) == 1:
        cutoffs = [0, cutoffs[0]]
    elif len(cutoffs) != 2:
        raise ValueError("Expected one value or two values separated by comma for the quality cutoff")
    return cutoffs


@dataclass
class TrimParams:
    """The hot-path subset of miRge's ``args`` namespace (SURVEY.md section 5 / parse.py)."""

    adapters: List[Adapter] = field(default_factory=list)
    times: int = 1  # parse.py:96
    nextseq_trim: Optional[int] = None  # parse.py:79
    quality_cutoff: Optional[str] = "10"  # parse.py:80 (always on by default)
    quality_base: int = 33  # parse.py:40 ("phred64", really the base)
    trim_n: bool = False  # parse.py:82
    cut: List[int] = field(default_factory=list)  # parse.py:78
    minimum_length: int = 16  # parse.py:83
    umi: Optional[Tuple[int, int]] = None  # parse.py:84 "-umi f,b"
    qiagenumi: bool = False  # parse.py:85
    count_mode: str = "head"  # "head": digest.py:354-373 as written; "release": dist/ 0.1.x
    compat: str = "2-3"  # cutadapt's alignment objective: "2-3" (matches, then cost) or "4" (score, then cost)

    def modifiers(self) -> List[Tuple[str, tuple]]:
        """Ordered modifier list exactly as stipulate() builds it (digest.py:87-99):
        NextSeq -> Quality -> AdapterCutter -> NEnd -> UnconditionalCutter(s)."""
        mods: List[Tuple[str, tuple]] = []
        if self.nextseq_trim is not None:
            mods.append(("nextseq", (int(self.nextseq_trim), self.quality_base)))
        if self.quality_cutoff is not None:
            q5, q3 = parse_cutoffs(self.quality_cutoff)
            mods.append(("quality", (q5, q3, self.quality_base)))
        if self.adapters:
            mods.append(("adapter", ()))
        if self.trim_n:
            mods.append(("nend", ()))
        cut = [c for c in self.cut]
        if cut:
            if len(cut) > 2:
                raise ValueError("You cannot remove bases from more than two ends.")
            if len(cut) == 2 and cut[0] * cut[1] > 0:
                raise ValueError("You cannot remove bases from the same end twice.")
            for c in cut:
                if c != 0:
                    mods.append(("cut", (int(c),)))
        return mods


def apply_modifier(mod, seq: str, qual: str, start: int, stop: int, p: TrimParams) -> Tuple[int, int]:
    """Apply one modifier to the window read[start:stop]; returns the new window."""
    kind, a = mod
    cur_s = seq[start:stop]
    cur_q = qual[start:stop]
    if kind == "nextseq":
        return start, start + nextseq_trim_index(cur_s, cur_q, a[0], a[1])
    if kind == "quality":
        s, e = quality_trim_index(cur_q, a[0], a[1], a[2])
        return start + s, start + e
    if kind == "adapter":
        for _ in range(p.times):
            ad, mt = best_match(p.adapters, seq[start:stop], p.compat)
            if mt is None:
                break
            if ad.where == "linked":
                start, stop = start + mt[2], start + mt[3]  # LinkedMatch.trimmed(): both ends cut
            elif ad.where in REMOVE_BEFORE:
                start = start + mt[3]  # read[rstop:]
            else:
                stop = start + mt[2]  # read[:rstart]
        return start, stop
    if kind == "nend":
        # NEndTrimmer: regex ^N+ and N+$ (uppercase N only)
        s, e = start, stop
        while s < e and seq[s] == "N":
            s += 1
        while e > s and seq[e - 1] == "N":
            e -= 1
        return s, e
    if kind == "cut":
        c = a[0]
        ln = stop - start
        if c > 0:
            return start + min(c, ln), stop  # read[c:]
        return start, start + max(ln + c, 0)  # read[:c], c < 0
    raise ValueError(kind)


def umi_parser(s: str, f: int, b: int) -> Tuple[str, str]:
    """UMIParser (digest.py:305-315) incl. the b == 0 quirk (s[-0:] is the whole string)."""
    front = s[:f]
    center = s[f:-b] if int(b) != 0 else s[f:]
    end = s[-b:]
    return center, front + end


def digest_read(seq: str, qual: str, p: TrimParams) -> List[Tuple[str, List[Tuple[int, int, int, int]]]]:
    """Per-read body of the worker ``cutadapt(n)`` (digest.py:325-373).

    Returns the list of emitted keys for this read, each with its window description
    (start, stop, ustart, ustop): key == seq[start:stop] + seq[ustart:ustop].
    """
    mods = p.modifiers()
    out = []
    start, stop = 0, len(seq)
    min_len = int(p.minimum_length)
    if p.qiagenumi:
        # digest.py:332-352 -- whole pipeline first, then UMI = text after the trimmed read
        for mod in mods:
            start, stop = apply_modifier(mod, seq, qual, start, stop, p)
        trimmed = seq[start:stop]
        U = int(p.umi[1])
        m_ad = len(p.adapters[0].sequence)  # qiaAdapter = args.adapters[0][1] (digest.py:121)
        ustart = ustop = 0
        if trimmed == "":
            umi_seq = ""  # "".split("") raises ValueError -> umi_seq = "" (digest.py:345-346)
        else:
            first = seq.find(trimmed)
            after = first + len(trimmed)
            nxt = seq.find(trimmed, after)
            seg_end = len(seq) if nxt < 0 else nxt
            # umi_seq = seg[:max_ad][-U:]
            seg_end = min(seg_end, after + m_ad + U)
            ustop = seg_end
            ustart = max(after, seg_end - U) if U != 0 else after  # [-0:] keeps the whole string
            umi_seq = seq[after:seg_end]
            umi_seq = umi_seq[-U:] if U != 0 else umi_seq
            assert umi_seq == seq[ustart:ustop]
            assert seq.split(trimmed)[1][: m_ad + U][-U:] == umi_seq if U != 0 else True
        if len(trimmed) >= min_len:
            out.append((trimmed + umi_seq, (start, stop, ustart, ustop)))
        return out
    emit_each = p.count_mode == "head"
    for si, mod in enumerate(mods):
        start, stop = apply_modifier(mod, seq, qual, start, stop, p)
        if emit_each or si == len(mods) - 1:
            cur = seq[start:stop]
            if p.umi is not None:
                ln = len(umi_parser(cur, p.umi[0], p.umi[1])[0])
            else:
                ln = len(cur)
            if ln >= min_len:
                out.append((cur, (start, stop, 0, 0)))
    return out


def digest_chunk(data: bytes, p: TrimParams) -> Tuple[int, Dict[str, int]]:
    """The worker ``cutadapt(n)`` (digest.py:320-375): (records parsed, chunk-local dict)."""
    read_dict: Dict[str, int] = {}
    count = 0
    for _name, seq, qual in parse_fastq(data):
        count += 1
        for key, _w in digest_read(seq, qual, p):
            read_dict[key] = read_dict.get(key, 0) + 1
    return count, read_dict


@dataclass
class SampleDigest:
    count: int  # sampleReadCounts (digest.py:214)
    trimmed: int  # trimmedReadCounts (digest.py:183/204/208)
    table: Dict[str, int]  # completeDict after the optional UMI level (digest.py:182/203)
    rlen: List[int]  # visual_treat['rlen'] (digest.py:146-157): one length per chunk-unique key
    hist: List[int]  # visual_treat['hist'] (digest.py:172/192)
    umi_rows: List[Tuple[str, str, int]]  # <sample>_umiCounts.csv rows (digest.py:196)


def digest_sample(data: bytes, p: TrimParams, umi_dedup: bool = False, buffer_size: int = 4_000_000) -> SampleDigest:
    """One iteration of baking()'s per-sample loop (digest.py:133-217)."""
    count = trimmed = 0
    complete: Dict[str, int] = {}
    rlen: List[int] = []
    for s, e in read_chunks(data, buffer_size):
        a, b = digest_chunk(data[s:e], p)
        count += a
        for k, c in b.items():
            if p.umi is not None:
                rlen.append(len(umi_parser(k, p.umi[0], p.umi[1])[0]))
            else:
                rlen.append(len(k))
            complete[k] = complete.get(k, 0) + c
            trimmed += c
    hist: List[int] = []
    umi_rows: List[Tuple[str, str, int]] = []
    if p.umi is not None:
        trimmed = 0
        second: Dict[str, int] = {}
        for s_, c in complete.items():
            pure, cut = umi_parser(s_, p.umi[0], p.umi[1])
            hist.append(c)
            if len(pure) >= int(p.minimum_length):
                if umi_dedup:
                    umi_rows.append((cut, pure, c))
                    second[pure] = second.get(pure, 0) + 1
                    trimmed += 1
                else:
                    second[pure] = second.get(pure, 0) + c
                    trimmed += c
        complete = second
    return SampleDigest(count, trimmed, complete, rlen, hist, umi_rows)


# --------------------------------------------------------------------------------------
# Annotation rounds  (bwtAlign, manifoldAlign.py:68-146; bowtie 1.x semantics)
# --------------------------------------------------------------------------------------

ROUND_COLUMNS = ["exact miRNA", "hairpin miRNA", "mature tRNA", "primary tRNA", "snoRNA", "rRNA",
                 "ncrna others", "mRNA", "isomiR miRNA", "spike-in"]  # digest.py:253
ROUND_LIBS = ["mirna", "hairpin", "mature_trna", "pre_trna", "snorna", "rrna", "ncrna_others", "mrna",
              "mirna", "spike-in"]  # manifoldAlign.py:84


@dataclass(frozen=True)
class RoundPolicy:
    """Effective bowtie policy of one round (manifoldAlign.py:85).  FASTA input => every quality
    is 'I' (Phred 40, Maq-rounded to 30); ``-e 70`` => at most 2 mismatches in total in -n mode."""

    seed_len: int  # -l 28 in -n mode; 0 => whole read (-v mode)
    seed_mm: int  # -n N / -v N
    total_mm: int  # 2 in -n mode (floor(70/30)); N in -v mode
    trim5: int = 0  # -5
    trim3: int = 0  # -3
    strip_polyT: bool = False  # round 3 query rewrite (manifoldAlign.py:118-126)


def _n(nmm):
    return RoundPolicy(28, nmm, 2)


ROUND_POLICIES = [
    _n(0),  # 0: -n 0
    _n(1),  # 1: -n 1
    RoundPolicy(0, 1, 1),  # 2: -v 1 -a --best --strata
    RoundPolicy(0, 0, 0, strip_polyT=True),  # 3: -v 0 -a --best --strata
    _n(1), _n(1), _n(1),  # 4,5,6: -n 1
    _n(0),  # 7: -n 0
    RoundPolicy(0, 2, 2, trim5=1, trim3=2),  # 8: -5 1 -3 2 -v 2 --best
    _n(0),  # 9: -n 0 (spike-in, only with -spk)
]

_POLYT = re.compile("T{3,}$")


@dataclass
class Library:
    names: List[str]  # FASTA header up to first whitespace (bowtie RNAME)
    seqs: List[str]


def read_fasta(text: str) -> Library:
    names, seqs, cur = [], [], None
    for line in text.splitlines():
        if line.startswith(">"):
            hdr = line[1:].split()
            names.append(hdr[0] if hdr else "")
            seqs.append([])
            cur = seqs[-1]
        elif cur is not None:
            cur.append(line.strip())
    return Library(names, ["".join(s).upper() for s in seqs])


def round_query(seq: str, rnd: int) -> Optional[str]:
    """Query string bowtie sees for ``seq`` in round ``rnd`` (None => not submitted)."""
    pol = ROUND_POLICIES[rnd]
    if pol.strip_polyT:
        mt = _POLYT.search(seq)
        if mt is None:
            return None
        q = seq[: mt.span(0)[0]]
    else:
        q = seq
    if pol.trim5 or pol.trim3:
        q = q[pol.trim5 : max(len(q) - pol.trim3, pol.trim5)]
    return q


def hits(query: str, lib: Library, pol: RoundPolicy) -> List[Tuple[int, int, int, int]]:
    """All valid end-to-end ungapped forward-strand alignments (n_mismatch, ref, offset, seed_mm).
    A read character other than ACGT always mismatches; a reference position other than ACGT may
    not be overlapped."""
    q = query.upper()
    L = len(q)
    out = []
    if L == 0:
        return out
    seed = L if pol.seed_len == 0 else min(pol.seed_len, L)
    for r, ref in enumerate(lib.seqs):
        for off in range(0, len(ref) - L + 1):
            mm = smm = 0
            ok = True
            for j in range(L):
                rc = ref[off + j]
                if rc not in "ACGT":
                    ok = False
                    break
                if q[j] != rc or q[j] not in "ACGT":
                    mm += 1
                    if j < seed:
                        smm += 1
                    if mm > pol.total_mm or smm > pol.seed_mm:
                        ok = False
                        break
            if ok:
                out.append((mm, r, off, smm))
    return out


def canonical_pick(hitlist):
    """min over (n_mismatch, ref index in library order, 0-based offset) -- the documented
    replacement for bowtie's irreproducible choice."""
    return min((h[0], h[1], h[2]) for h in hitlist) if hitlist else None


def annotate(seqs: Sequence[str], libs: Dict[str, Library], spike_in: bool = False):
    """bwtAlign's round loop (manifoldAlign.py:90-135) over the unique sequences.

    Returns {seq: (round, ref_name, offset, n_mismatch)} for annotated sequences; a sequence is
    annotated by the first round that hits it (rounds 0 and 1 partition by length and ignore
    annotFlag, manifoldAlign.py:93,104)."""
    annot: Dict[str, Tuple[int, str, int, int]] = {}
    rounds = 10 if spike_in else 9
    for rnd in range(rounds):
        lib = libs[ROUND_LIBS[rnd]]
        pol = ROUND_POLICIES[rnd]
        for s in seqs:
            if rnd == 0:
                if not len(s) < 26:
                    continue
            elif rnd == 1:
                if not len(s) > 25:
                    continue
            elif s in annot:
                continue
            q = round_query(s, rnd)
            if q is None:
                continue
            pick = canonical_pick(hits(q, lib, pol))
            if pick is not None:
                annot[s] = (rnd, lib.names[pick[1]], pick[2], pick[0])
    return annot


# ---------------------------------------------------------------------------------------------------
# Split + summarize counters: mirge/__main__.py:164-173 and the parts of
# mirge/libs/summary.py::summarize that feed annotation.report.csv and miR.Counts.csv.  Pure Python over
# dicts; pinned byte-for-byte on files the unmodified reference wrote (tests/golden/ref_case1, generated by
# tests/golden/make_reference_golden.py).
# ---------------------------------------------------------------------------------------------------

REPORT_COLUMNS = ["Total Input Reads", "Trimmed Reads (all)", "Trimmed Reads (unique)", "All miRNA Reads", "Filtered miRNA Reads",
                  "Unique miRNAs", "Hairpin miRNAs", "mature tRNA Reads", "primary tRNA Reads", "snoRNA Reads", "rRNA Reads",
                  "ncRNA others", "mRNA Reads", "Spike-in", "Remaining Reads"]  # summary.py:897,901 (colRearrange)
_REPORT_ROUND = {"Hairpin miRNAs": 1, "mature tRNA Reads": 2, "primary tRNA Reads": 3, "snoRNA Reads": 4, "rRNA Reads": 5,
                 "ncRNA others": 6, "mRNA Reads": 7, "Spike-in": 9}  # summary.py:686-691 col_headers -> round


def read_merges(text: str) -> Tuple[Dict[str, str], List[str]]:
    """<organism>_merges_<db>.csv (summary.py:705-714): member name -> merged name, and the merged names."""
    member, merged = {}, []
    for line in text.splitlines():
        c = line.strip().split(",")
        for item in c[1:]:
            member[item] = c[0]
        merged.append(c[0])
    return member, merged


def table_csv(rows: Sequence[Tuple[str, int, Sequence[str], Sequence[int]]], samples: Sequence[str], spike_in: bool) -> str:
    """``DataFrame.to_csv`` of mapped.csv / unmapped.csv (__main__.py:172-173): index label Sequence, annotFlag,
    the annotation columns ('spike-in' only with -spk, manifoldAlign.py:137-138), one column per sample."""
    cols = ROUND_COLUMNS if spike_in else ROUND_COLUMNS[:9]
    out = ["Sequence,annotFlag," + ",".join(cols) + "," + ",".join(samples)]
    for seq, flag, names, counts in rows:
        out.append("%s,%d,%s,%s" % (seq, flag, ",".join(names[: len(cols)]), ",".join(str(int(c)) for c in counts)))
    return "\n".join(out) + "\n"


def split_tables(annot: Dict[str, Tuple[int, str, int, int]], counts: Dict[str, Sequence[int]], samples: Sequence[str],
                 spike_in: bool) -> Tuple[str, str]:
    """(mapped.csv, unmapped.csv) texts: rows in the DataFrame's order (sorted sequences), annotFlag split
    of __main__.py:164-165."""
    mapped, unmapped = [], []
    for seq in sorted(counts):
        names = [""] * 10
        a = annot.get(seq)
        if a is not None:
            names[a[0]] = a[1]
            mapped.append((seq, 1, names, counts[seq]))
        else:
            unmapped.append((seq, 0, names, counts[seq]))
    return table_csv(mapped, samples, spike_in), table_csv(unmapped, samples, spike_in)


def canonical_filter(can: int, iso: int, ca_thr: float) -> int:
    """mirge_can (summary.py:25-45) for one miRNA of one sample: exact reads ``can``, isomiR reads ``iso``."""
    if can < 2:  # :36-37
        can, iso = 0, 0
    ratio = can / iso if iso > 0 else float(can)  # :38-39
    return can + iso if ratio > ca_thr else 0  # :41-42


def summarize_counts(annot: Dict[str, Tuple[int, str, int, int]], counts: Dict[str, Sequence[int]], samples: Sequence[str],
                     sample_reads: Dict[str, int], trimmed: Dict[str, int], trimmed_unique: Dict[str, int], merges_text: str,
                     mirna_names: Sequence[str], ca_thr: float = 0.1, spike_in: bool = False):
    """annotation.report.csv rows and miR.Counts.csv rows as summarize() computes them.
    Returns (report {sample: {column: int}}, mir_counts {merged miRNA name: [float per sample]})."""
    S = len(samples)
    member, merged_names = read_merges(merges_text)
    lib_sum = {k: [0] * S for k in _REPORT_ROUND}
    all_mir = [0] * S
    can: Dict[str, List[int]] = {}
    iso: Dict[str, List[int]] = {}
    for seq, (rnd, name, _off, _mm) in annot.items():
        c = counts[seq]
        for col, r in _REPORT_ROUND.items():  # summary.py:692-698
            if r == rnd:
                for j in range(S):
                    lib_sum[col][j] += c[j]
        if rnd in (0, 8):  # summary.py:720,764-766 ('All miRNA Reads')
            tgt = can if rnd == 0 else iso
            row = tgt.setdefault(name, [0] * S)
            for j in range(S):
                row[j] += c[j]
                all_mir[j] += c[j]
    grouped: Dict[str, List[float]] = {}
    for name in sorted(can):  # cann_collapse rows (groupby sorts), left-merged with iso_collapse (summary.py:741-748)
        y = iso.get(name, [0] * S)
        out_name = member.get(name, name)  # :750-752
        row = grouped.setdefault(out_name, [0.0] * S)
        for j in range(S):
            row[j] += float(canonical_filter(can[name][j], y[j], ca_thr))
    report = {}
    for j, s in enumerate(samples):
        r = {"Total Input Reads": sample_reads[s], "Trimmed Reads (all)": trimmed[s], "Trimmed Reads (unique)": trimmed_unique[s],
             "All miRNA Reads": all_mir[j], "Filtered miRNA Reads": int(sum(v[j] for v in grouped.values())),
             "Unique miRNAs": sum(1 for v in grouped.values() if v[j] > 0)}  # :757-758, :882-887
        for col in _REPORT_ROUND:
            if col == "Spike-in" and not spike_in:
                continue
            r[col] = lib_sum[col][j]
        tosum = ["All miRNA Reads"] + [c for c in _REPORT_ROUND if c in r]  # :896,900 col_tosum
        r["Remaining Reads"] = r["Trimmed Reads (all)"] - sum(r[c] for c in tosum)  # :1224
        report[s] = r
    # miR.Counts.csv: every library miRNA (bowtie-inspect -n, 'segs:' names cut at the first blank) that is not a
    # member of a merge, plus the merged names, outer-joined with the grouped counts (summary.py:774-797)
    names = list(merged_names)
    for srow in mirna_names:
        if "segs:" in srow:
            srow = srow.split(" ")[0]
        if srow not in member:
            names.append(srow)
    mir_counts = {n: grouped.get(n, [0.0] * S) for n in sorted(set(names) | set(grouped))}
    return report, mir_counts


def report_csv(report: Dict[str, Dict[str, int]], samples: Sequence[str], spike_in: bool) -> str:
    cols = [c for c in REPORT_COLUMNS if spike_in or c != "Spike-in"]
    out = ["Sample name(s)," + ",".join(cols)]
    for s in samples:
        out.append(s + "," + ",".join(str(int(report[s][c])) for c in cols))
    return "\n".join(out) + "\n"


def mir_counts_csv(mir_counts: Dict[str, Sequence[float]], samples: Sequence[str]) -> str:
    out = ["miRNA," + ",".join(samples)]
    for n, v in mir_counts.items():
        out.append(n + "," + ",".join(repr(float(x)) for x in v))
    return "\n".join(out) + "\n"


# ---------------------------------------------------------------------------------------------------
# bowtie 1 SAM records (-S) for the per-round SAM files of -bam / -trf (manifoldAlign.py:20-62).  Field layout
# after the bowtie manual and the example line the reference quotes (summary.py:1194): FLAG 0 (forward strand,
# --norc), 1-based POS of the (trimmed) query, MAPQ 255, CIGAR <len>M, QUAL all 'I' (FASTA input), then
# XA:i:<stratum> MD:Z:<mismatch string> NM:i:<mismatches>.  bowtie's XM:i tag is not reproduced (its value
# depends on -m/-M bookkeeping the reference never reads).
# ---------------------------------------------------------------------------------------------------

def md_string(query: str, ref_seg: str) -> Tuple[str, List[int]]:
    """MD:Z value of an ungapped alignment and the 0-based query positions of its mismatches."""
    out, run, pos = [], 0, []
    for j, (q, r) in enumerate(zip(query.upper(), ref_seg.upper())):
        if q == r and q in "ACGT":
            run += 1
        else:
            out.append(str(run))
            out.append(r)
            run = 0
            pos.append(j)
    out.append(str(run))
    return "".join(out), pos


def sam_line(qname: str, query: str, ref_name: str, ref_seq: str, off: int, pol: RoundPolicy) -> str:
    md, pos = md_string(query, ref_seq[off : off + len(query)])
    seed = len(query) if pol.seed_len == 0 else min(pol.seed_len, len(query))
    stratum = sum(1 for p in pos if p < seed)
    return "%s\t0\t%s\t%d\t255\t%dM\t*\t0\t0\t%s\t%s\tXA:i:%d\tMD:Z:%s\tNM:i:%d" % (
        qname, ref_name, off + 1, len(query), query, "I" * len(query), stratum, md, len(pos))


def mir_rpm_csv(mir_counts: Dict[str, Sequence[float]], samples: Sequence[str]) -> str:
    """miR.RPM.csv (summary.py:759,795,797): counts / column sum * 1e6 rounded to 4 decimals (numpy rounding, as
    pandas does), names without reads stay 0.0."""
    import numpy as np

    names = list(mir_counts)
    m = np.array([list(mir_counts[n]) for n in names], dtype=np.float64).reshape(len(names), len(samples))
    with np.errstate(divide="ignore", invalid="ignore"):
        rpm = np.nan_to_num(np.round(m / m.sum(axis=0) * 1000000, 4), nan=0.0)
    out = ["miRNA," + ",".join(samples)]
    for n, row in zip(names, rpm.tolist()):
        out.append(n + "," + ",".join(repr(float(x)) for x in row))
    return "\n".join(out) + "\n"
