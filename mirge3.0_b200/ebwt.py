"""Decoder for bowtie 1 index files (``<base>.1.ebwt`` / ``.3.ebwt`` / ``.4.ebwt``): recovers (names, sequences)
so that a stock ``miRge3_Lib`` -- which ships only bowtie indexes (manifoldAlign.py:97-98) -- can be loaded without
``bowtie-inspect`` (summary.py:776-788,812-827; bamFmt.py:9-33).  SURVEY.md section 8f row 1 / Appendix D.

Only three pieces of the index are needed, none of them the BWT:
  ``.3.ebwt``  int32 1 (endianness), uint32 n, then n records {uint32 off, uint32 len, uint8 first}: one per
               unambiguous stretch; ``off`` ambiguous characters precede it, ``first`` starts a new reference;
  ``.4.ebwt``  the unambiguous bases, 2 bits each (A=0 C=1 G=2 T=3), base i in bits 2*(i%4) of byte i/4;
  ``.1.ebwt``  after the header and the tables, the reference names as newline-separated text closed by a NUL.

Like ``bowtie-inspect``, the result has ``N`` for every ambiguous character inside a reference and loses
trailing ambiguous characters (the index does not record them).

STATUS: written from the published layout of bowtie 1.x (ebwt.h ``EbwtParams`` / ``Ebwt::readIntoMemory``,
``RefRecord``), not yet checked against an index built by a real bowtie-build (none is available here).  The
reader therefore validates what it can and refuses instead of guessing: the name block located from the header
arithmetic must end exactly at the file's closing NUL and hold one printable name per reference, the ``.3`` record
lengths must add up to the ``len`` of the ``.1`` header and fit the ``.4`` file."""
from __future__ import annotations

import os
import struct
from typing import List, Tuple

import numpy as np

from .device import MirgeError


def _u32(buf, off):
    return struct.unpack_from("<I", buf, off)[0]


def read_records(path3: str):
    """(off[], len[], first[]) of ``<base>.3.ebwt``."""
    raw = open(path3, "rb").read()
    if len(raw) < 8:
        raise MirgeError("%s: truncated" % path3)
    one = struct.unpack_from("<i", raw, 0)[0]
    if one != 1:
        raise MirgeError("%s: not a little-endian bowtie 1 index (large .ebwtl indexes are not supported)" % path3)
    n = _u32(raw, 4)
    if 8 + 9 * n > len(raw):
        raise MirgeError("%s: %d records do not fit the file" % (path3, n))
    rec = np.frombuffer(raw, dtype=np.uint8, count=9 * n, offset=8).reshape(n, 9)
    off = rec[:, 0:4].copy().view("<u4").reshape(n).astype(np.int64)
    ln = rec[:, 4:8].copy().view("<u4").reshape(n).astype(np.int64)
    first = rec[:, 8] != 0
    return off, ln, first


def names_offset(raw: bytes) -> int:
    """Offset of the reference names inside ``<base>.1.ebwt`` from the header arithmetic (EbwtParams)."""
    one, length, line_rate, lines_per_side, off_rate, ftab_chars, flags = struct.unpack_from("<7i", raw, 0)
    if one != 1:
        raise MirgeError("not a little-endian bowtie 1 index")
    length &= 0xFFFFFFFF
    pos = 28
    n_pat = _u32(raw, pos)
    pos += 4 + 4 * n_pat  # plen[]
    n_frag = _u32(raw, pos)
    pos += 4 + 12 * n_frag  # rstarts[]: 3 words per fragment
    bwt_sz = length // 4 + 1
    side_sz = (1 << line_rate) * lines_per_side
    side_bwt_sz = side_sz - 8
    n_side_pairs = (bwt_sz + 2 * side_bwt_sz - 1) // (2 * side_bwt_sz)
    pos += n_side_pairs * 2 * side_sz  # ebwt[]
    pos += 4 + 5 * 4  # zOff, fchr[5]
    pos += 4 * ((1 << (2 * ftab_chars)) + 1)  # ftab[]
    pos += 4 * (2 * ftab_chars)  # eftab[]
    return pos


def _printable(b: bytes) -> bool:
    return all(32 <= c < 127 or c == 9 for c in b)


def read_names(path1: str, n_refs: int) -> Tuple[List[str], int]:
    """Reference names (full header lines, as ``bowtie-inspect -n`` prints them) and the header's ``len``."""
    raw = open(path1, "rb").read()
    if len(raw) < 32:
        raise MirgeError("%s: truncated" % path1)
    length = struct.unpack_from("<i", raw, 4)[0] & 0xFFFFFFFF
    end = len(raw)
    if raw[end - 1] != 0:
        raise MirgeError("%s: the name block is not closed by a NUL byte" % path1)
    names = None
    try:
        pos = names_offset(raw)
        if 0 < pos < end:
            cand = raw[pos : end - 1].split(b"\n")
            if cand and cand[-1] == b"":
                cand = cand[:-1]
            if len(cand) == n_refs and all(_printable(c) for c in cand):
                names = cand
    except (struct.error, MirgeError):
        names = None
    if names is None:
        # layout differs from the header arithmetic: take the longest printable tail, which starts right after
        # the binary eftab[] words, and insist on exactly one line per reference
        p = end - 1
        while p > 0 and (32 <= raw[p - 1] < 127 or raw[p - 1] in (9, 10)):
            p -= 1
        cand = raw[p : end - 1].split(b"\n")
        if cand and cand[-1] == b"":
            cand = cand[:-1]
        if len(cand) != n_refs:
            raise MirgeError("%s: cannot locate the %d reference names (found %d lines at the end of the file)"
                             % (path1, n_refs, len(cand)))
        names = cand
    return [c.decode("latin-1") for c in names], length


def decode_index(basename: str) -> Tuple[List[str], List[bytes]]:
    """(names, sequences) of the bowtie 1 index ``basename`` -- the output of ``bowtie-inspect -a`` without
    running it.  ``names`` are cut at the first whitespace (bowtie's RNAME); ``decode_names`` keeps the whole
    header line."""
    full, seqs = decode_full(basename)
    return [n.split()[0] if n.split() else "" for n in full], seqs


def decode_full(basename: str) -> Tuple[List[str], List[bytes]]:
    p1, p3, p4 = basename + ".1.ebwt", basename + ".3.ebwt", basename + ".4.ebwt"
    for p in (p1, p3, p4):
        if not os.path.exists(p):
            raise MirgeError("bowtie index file %s is missing" % p)
    off, ln, first = read_records(p3)
    if off.size == 0 or not first[0]:
        raise MirgeError("%s: the first record does not start a reference" % p3)
    n_refs = int(first.sum())
    names, length = read_names(p1, n_refs)
    total = int(ln.sum())
    if total != length:
        raise MirgeError("%s: stretch lengths add up to %d, the .1.ebwt header says %d" % (p3, total, length))
    packed = np.fromfile(p4, dtype=np.uint8)
    if packed.size * 4 < total:
        raise MirgeError("%s: %d bases expected, file holds %d" % (p4, total, packed.size * 4))
    codes = ((packed[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(-1)[:total]
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    seqs: List[bytes] = []
    cur: List[bytes] = []
    pos = 0
    for o, l, f in zip(off.tolist(), ln.tolist(), first.tolist()):
        if f and (cur or seqs):
            seqs.append(b"".join(cur))
            cur = []
        if o:
            cur.append(b"N" * o)
        cur.append(text[pos : pos + l].tobytes())
        pos += l
    seqs.append(b"".join(cur))
    if len(seqs) != n_refs:
        raise MirgeError("%s: %d references reconstructed, %d expected" % (basename, len(seqs), n_refs))
    return names, seqs


def decode_names(basename: str) -> List[str]:
    """``bowtie-inspect -n <basename>``: the full name line of every reference (summary.py:776-788)."""
    _, _, first = read_records(basename + ".3.ebwt")
    return read_names(basename + ".1.ebwt", int(first.sum()))[0]
