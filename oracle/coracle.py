"""ctypes loader for the C oracle (oracle/mirge_oracle.c).  TEST INFRASTRUCTURE ONLY -- see the
header of pyoracle.py.  PARITY UNPINNED (no cutadapt/bowtie available to pin against)."""
import ctypes as C
import os
import subprocess

import numpy as np

import mirge_b200  # noqa: F401  (import shim for the product package's ABI structs)
from mirge_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libmirge_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "mirge_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "mirge_b200.h")
    if force or not os.path.exists(SO) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(SO) for f in (src, hdr)
    ):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "_build/libmirge_oracle.so"])
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        P, U64 = C.c_void_p, C.c_uint64
        L.oracle_line_index.restype = C.c_int64
        L.oracle_line_index.argtypes = [P, U64, P, U64]
        L.oracle_trim_slots.restype = C.c_int
        L.oracle_trim_slots.argtypes = [C.POINTER(abi.TrimParams)]
        L.oracle_trim.restype = C.c_int64
        L.oracle_trim.argtypes = [P, U64, C.POINTER(abi.TrimParams), P, U64, P, P, C.c_int]
        L.oracle_digest_collapse.restype = P
        L.oracle_digest_collapse.argtypes = [P, U64, C.POINTER(abi.TrimParams), P, U64, C.c_int, C.POINTER(C.c_int64)]
        L.oracle_umi_collapse.restype = P
        L.oracle_umi_collapse.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int]
        for n in ("oracle_table_size", "oracle_table_bytes", "oracle_table_total"):
            getattr(L, n).restype = U64
            getattr(L, n).argtypes = [P]
        L.oracle_table_export.restype = None
        L.oracle_table_export.argtypes = [P, P, P, P]
        L.oracle_table_free.restype = None
        L.oracle_table_free.argtypes = [P]
        L.oracle_annotate_round.restype = None
        L.oracle_annotate_round.argtypes = [P, P, U64, P, P, C.c_uint32, C.POINTER(abi.RoundPolicy), P, P, C.c_int]
        L.oracle_index_build.restype = P
        L.oracle_index_build.argtypes = [P, P, C.c_uint32, C.c_int]
        L.oracle_index_free.restype = None
        L.oracle_index_free.argtypes = [P]
        L.oracle_annotate_round_indexed.restype = None
        L.oracle_annotate_round_indexed.argtypes = [P, P, U64, P, C.POINTER(abi.RoundPolicy), P, P, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def line_index(fq: np.ndarray):
    """(n_records, line_start u32[4n+1]); raises ValueError on malformed input."""
    L = lib()
    n = L.oracle_line_index(_ptr(fq), fq.size, None, 0)
    if n < 0:
        raise ValueError("FASTQ format error (line count not a multiple of 4)")
    ls = np.zeros(4 * n + 1, dtype=np.uint32)
    L.oracle_line_index(_ptr(fq), fq.size, _ptr(ls), n)
    return int(n), ls


def trim(fq: np.ndarray, params: abi.TrimParams, nthreads: int = 1):
    """(n_records, win u16[n, E, 4], kept u8[n, E]) for every record of the FASTQ bytes."""
    L = lib()
    n, ls = line_index(fq)
    E = L.oracle_trim_slots(C.byref(params))
    win = np.zeros((n, E, 4), dtype=np.uint16)
    kept = np.zeros((n, E), dtype=np.uint8)
    rc = L.oracle_trim(_ptr(fq), fq.size, C.byref(params), _ptr(ls), n, _ptr(win), _ptr(kept), nthreads)
    if rc < 0:
        raise ValueError("FASTQ format error")
    return n, win, kept


class Table:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        try:
            if self.h:
                lib().oracle_table_free(self.h)
                self.h = None
        except Exception:  # interpreter shutdown
            pass

    @property
    def total(self):
        return int(lib().oracle_table_total(self.h))

    def __len__(self):
        return int(lib().oracle_table_size(self.h))

    def export(self):
        """(keys bytes, key_off u64[n+1], counts u64[n])"""
        L = lib()
        n = len(self)
        keys = np.zeros(max(1, int(L.oracle_table_bytes(self.h))), dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.uint64)
        cnt = np.zeros(n, dtype=np.uint64)
        L.oracle_table_export(self.h, _ptr(keys), _ptr(off), _ptr(cnt))
        return keys, off, cnt

    def to_dict(self):
        keys, off, cnt = self.export()
        b = keys.tobytes()
        return {b[int(off[i]) : int(off[i + 1])].decode("latin-1"): int(cnt[i]) for i in range(len(cnt))}

    def umi_collapse(self, f, b, min_len, dedup):
        return Table(lib().oracle_umi_collapse(self.h, f, b, min_len, 1 if dedup else 0))


def digest_collapse(fq: np.ndarray, params: abi.TrimParams, nthreads: int = 1):
    """(n_records, Table) -- completeDict of one sample before the UMI level."""
    L = lib()
    n, ls = line_index(fq)
    st = C.c_int64(0)
    h = L.oracle_digest_collapse(_ptr(fq), fq.size, C.byref(params), _ptr(ls), n, nthreads, C.byref(st))
    t = Table(h)
    if st.value < 0:
        raise ValueError("FASTQ format error")
    return n, t


def annotate_round(keys: np.ndarray, key_off: np.ndarray, refs: np.ndarray, ref_off: np.ndarray,
                   policy: abi.RoundPolicy, annot_round: np.ndarray, hit: np.ndarray, nthreads: int = 1):
    lib().oracle_annotate_round(_ptr(keys), _ptr(key_off), len(key_off) - 1, _ptr(refs), _ptr(ref_off),
                                len(ref_off) - 1, C.byref(policy), _ptr(annot_round), _ptr(hit), nthreads)


class Index:
    """Sorted 16-mer index of one library for the indexed CPU search (keeps its inputs alive)."""

    def __init__(self, refs: np.ndarray, ref_off: np.ndarray):
        self.refs, self.ref_off = np.ascontiguousarray(refs), np.ascontiguousarray(ref_off, dtype=np.uint32)
        self.h = lib().oracle_index_build(_ptr(self.refs), _ptr(self.ref_off), len(self.ref_off) - 1, 1)

    def __del__(self):
        try:
            if self.h:
                lib().oracle_index_free(self.h)
                self.h = None
        except Exception:
            pass


def annotate_round_indexed(keys, key_off, index: Index, policy, annot_round, hit, nthreads=1):
    lib().oracle_annotate_round_indexed(_ptr(keys), _ptr(key_off), len(key_off) - 1, index.h, C.byref(policy),
                                        _ptr(annot_round), _ptr(hit), nthreads)
