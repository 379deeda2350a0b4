"""Host-side ingest of FASTQ files for ``baking`` -- what the reference does with ``xopen(FQfile, "rb")`` and
``dnaio.read_chunks`` in the parent process (mirge/libs/digest.py:136-140; SURVEY.md section 8, row a1 and "next" row f-3).

The GPU path consumes FASTQ bytes far faster than one core inflates them, and ``mirge_tokenise_sync`` makes the
digesting thread wait for the device once per piece, so reading and inflating must not happen on that thread:

* every file is read by a worker thread into a bounded queue of chunks; the digesting thread only copies chunks
  into its pinned staging buffers (``readinto``);
* plain gzip is one DEFLATE stream: large files are inflated chunk-parallel by the native decoder of csrc/pinflate.c
  (block starts searched near every cut, speculative decoding with an unknown window, chained and CRC-checked:
  libmirge_inflate.so), small ones -- or any file when that library has not been built -- by one worker with zlib
  (multi-member files and trailing zero padding are handled like Python's ``gzip`` module does either way);
* BGZF files (bgzip, many sequencing pipelines: gzip members of <= 64 KB that carry their size in a 'BC' extra field)
  are inflated block-parallel on a thread pool (zlib releases the GIL), CRC and size of every block checked;
* ``SampleReadahead`` keeps the readers of the next samples of a run going while the current one is digested, which
  is where a cohort of single-stream ``.fastq.gz`` files gets its parallelism (one inflating core per sample).

Nothing here touches the device; the classes are file-like sources for ``digest.HostStreamer``.  Truncated or corrupt
input raises (EOFError / OSError / zlib.error), like the reference's reader would."""
from __future__ import annotations

import collections
import ctypes
import mmap
import os
import queue
import struct
import threading
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Iterator, List, Optional, Sequence

CHUNK = 8 << 20  # bytes per queue entry (inflated)
RAW_READ = 1 << 20  # compressed bytes fed to zlib per call
BGZF_BATCH = 2 << 20  # compressed bytes per pool task
PGZ_MIN_BYTES = int(os.environ.get("MIRGE_B200_PARALLEL_GZIP_MIN_MB", "32")) << 20  # smaller gzip files: one zlib worker
PGZ_CHUNK = int(os.environ.get("MIRGE_B200_PARALLEL_GZIP_CHUNK_KB", "2048")) << 10  # compressed bytes per speculative chunk
PGZ_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmirge_inflate.so")

_pool_lock = threading.Lock()
_pool: Optional[ThreadPoolExecutor] = None
_pool_size = 0


def default_threads() -> int:
    return max(1, min(32, (os.cpu_count() or 2) - 1))


def inflate_pool(threads: Optional[int] = None) -> ThreadPoolExecutor:
    """Process-wide pool of inflate workers (grown, never shrunk)."""
    global _pool, _pool_size
    want = int(threads or default_threads())
    with _pool_lock:
        if _pool is None or want > _pool_size:
            _pool = ThreadPoolExecutor(max_workers=want, thread_name_prefix="mirge-inflate")
            _pool_size = want
        return _pool


def sniff(path: str) -> str:
    """'bgzf', 'gzip', 'bz2', 'xz' or 'plain' from the first bytes of the file (xopen decides by magic number too)."""
    with open(path, "rb") as f:
        head = f.read(18)
    if head[:3] == b"BZh":
        return "bz2"
    if head[:6] == b"\xfd7zXZ\x00":
        return "xz"
    if len(head) < 2 or head[:2] != b"\x1f\x8b":
        return "plain"
    if len(head) >= 18 and head[2] == 8 and (head[3] & 4):
        xlen = struct.unpack_from("<H", head, 10)[0]
        if xlen >= 6 and head[12:14] == b"BC" and struct.unpack_from("<H", head, 14)[0] == 2:
            return "bgzf"
    return "gzip"


# ---------------------------------------------------------------------------------------------------------------------
# producers: generators of inflated chunks, run on the reader thread


def _plain_chunks(path: str) -> Iterator[bytes]:
    with open(path, "rb", buffering=0) as f:
        while True:
            b = f.read(CHUNK)
            if not b:
                return
            yield b


def _gzip_chunks(path: str) -> Iterator[bytes]:
    """One worker, one DEFLATE stream at a time; members are chained, zero padding after a member is skipped."""
    with open(path, "rb", buffering=0) as f:
        d = zlib.decompressobj(31)
        fed = False  # the current member has received data
        while True:
            raw = f.read(RAW_READ)
            if not raw:
                if fed and not d.eof:
                    raise EOFError("Compressed file ended before the end-of-stream marker was reached: %s" % path)
                return
            buf = raw
            while buf:
                if not fed:  # between members: zero padding may run on into this read
                    buf = buf.lstrip(b"\x00")
                    if not buf:
                        break
                out = d.decompress(buf, CHUNK)
                fed = True
                if out:
                    yield out
                if d.eof:
                    buf = d.unused_data
                    d = zlib.decompressobj(31)
                    fed = False
                else:
                    buf = d.unconsumed_tail


def _module_chunks(opener, path: str) -> Iterator[bytes]:
    """bzip2 / xz input (xopen takes both): the standard library's reader, multi-stream files included, one core."""
    with opener(path, "rb") as f:
        while True:
            b = f.read(CHUNK)
            if not b:
                return
            yield b


_pgz = None


def parallel_gzip_library():
    """libmirge_inflate.so (csrc/pinflate.c, built by __graft_entry__.build()) or None when it has not been built or
    MIRGE_B200_PARALLEL_GZIP=0.  Host-only code: the serial zlib reader gives the same bytes without it."""
    global _pgz
    if os.environ.get("MIRGE_B200_PARALLEL_GZIP", "1") == "0":
        return None
    if _pgz is None:
        if not os.path.exists(PGZ_LIB):
            return None
        lib = ctypes.CDLL(PGZ_LIB)
        lib.pgz_open.restype = ctypes.c_void_p
        lib.pgz_open.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64]
        lib.pgz_read.restype = ctypes.c_int64
        lib.pgz_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        lib.pgz_error.restype = ctypes.c_char_p
        lib.pgz_error.argtypes = [ctypes.c_void_p]
        lib.pgz_close.argtypes = [ctypes.c_void_p]
        lib.pgz_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        _pgz = lib
    return _pgz


def _pgzip_chunks(path: str, threads: Optional[int] = None, chunk_bytes: Optional[int] = None,
                  spare: Optional[Callable[[], Optional[bytearray]]] = None) -> Iterator[bytes]:
    """One gzip stream inflated on ``threads`` cores (csrc/pinflate.c); same bytes and the same kinds of errors as
    _gzip_chunks: EOFError for a truncated file, OSError for CRC / length / header problems, zlib.error for bad data.
    ``spare()`` hands back a chunk buffer the consumer is done with (ChunkReader.spare), so that the steady state
    allocates, zero-fills and page-faults nothing."""
    lib = parallel_gzip_library()
    with open(path, "rb") as f:
        size = os.fstat(f.fileno()).st_size
        if size == 0:
            return
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    try:
        if hasattr(mm, "madvise") and hasattr(mmap, "MADV_SEQUENTIAL"):
            mm.madvise(mmap.MADV_SEQUENTIAL)
        addr = _address_of(mm)  # (no copy: the decoder reads the mapping)
        h = lib.pgz_open(addr, size, int(threads or default_threads()), int(chunk_bytes or PGZ_CHUNK))
        if not h:
            raise MemoryError("parallel gzip reader: out of memory")
        try:
            while True:
                out = spare() if spare is not None else None
                if out is None or len(out) != CHUNK:
                    out = bytearray(CHUNK)
                dst = (ctypes.c_ubyte * CHUNK).from_buffer(out)
                k = lib.pgz_read(h, dst, CHUNK)
                del dst
                if k < 0:
                    msg = lib.pgz_error(h).decode("latin-1") + ": " + path
                    if "ended before" in msg:
                        raise EOFError(msg)
                    if "invalid deflate" in msg:
                        raise zlib.error(msg)
                    raise OSError(msg)
                if k == 0:
                    return
                yield memoryview(out)[:k]
        finally:
            lib.pgz_close(h)
    finally:
        mm.close()


def _address_of(mm: "mmap.mmap") -> int:
    """address of a read-only mapping (ctypes' from_buffer wants a writable buffer; numpy takes any)"""
    import numpy as np

    return int(np.frombuffer(mm, dtype=np.uint8).ctypes.data)


def _read_exact(f, n: int, path: str) -> bytes:
    b = f.read(n)
    if len(b) != n:
        raise EOFError("Compressed file ended inside a BGZF block: %s" % path)
    return b


def _bgzf_blocks(f, path: str):
    """(deflate payload, crc32, isize) of every block, in file order."""
    while True:
        head = f.read(12)
        if not head:
            return
        if len(head) != 12 or head[:4] != b"\x1f\x8b\x08\x04":
            raise OSError("not a BGZF block at offset %d of %s" % (f.tell() - len(head), path))
        xlen = struct.unpack_from("<H", head, 10)[0]
        extra = _read_exact(f, xlen, path)
        bsize = None
        p = 0
        while p + 4 <= xlen:
            slen = struct.unpack_from("<H", extra, p + 2)[0]
            if extra[p : p + 2] == b"BC" and slen == 2:
                bsize = struct.unpack_from("<H", extra, p + 4)[0]
            p += 4 + slen
        if bsize is None:
            raise OSError("gzip member without a BGZF size field in %s" % path)
        rest = bsize + 1 - 12 - xlen
        if rest < 8:
            raise OSError("corrupt BGZF block size in %s" % path)
        body = _read_exact(f, rest, path)
        crc, isize = struct.unpack_from("<II", body, rest - 8)
        yield body[: rest - 8], crc, isize


def _inflate_blocks(blocks) -> bytes:
    out = []
    for payload, crc, isize in blocks:
        data = zlib.decompress(payload, -15) if isize else b""
        if len(data) != isize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
            raise OSError("BGZF block fails its CRC / size check")
        out.append(data)
    return b"".join(out)


def _bgzf_chunks(path: str, threads: Optional[int] = None) -> Iterator[bytes]:
    pool = inflate_pool(threads)
    window = 2 * (threads or default_threads())
    inflight: collections.deque = collections.deque()
    with open(path, "rb", buffering=4 << 20) as f:
        batch, size = [], 0
        for blk in _bgzf_blocks(f, path):
            batch.append(blk)
            size += len(blk[0])
            if size >= BGZF_BATCH:
                inflight.append(pool.submit(_inflate_blocks, batch))
                batch, size = [], 0
                while len(inflight) >= window:
                    out = inflight.popleft().result()
                    if out:
                        yield out
        if batch:
            inflight.append(pool.submit(_inflate_blocks, batch))
        while inflight:
            out = inflight.popleft().result()
            if out:
                yield out


# ---------------------------------------------------------------------------------------------------------------------


class ChunkReader:
    """File-like source (``readinto`` / ``read``) over chunks produced on a worker thread.

    The queue holds at most ``depth`` chunks, so a reader that is far ahead of its consumer blocks instead of
    buffering the file.  An exception on the worker is re-raised by the next ``readinto``."""

    def __init__(self, producer: Callable[[], Iterator[bytes]], depth: int = 4, name: str = ""):
        self.name = name
        self._q: queue.Queue = queue.Queue(maxsize=max(int(depth), 1))
        self._stop = threading.Event()
        self._cur = memoryview(b"")
        self._pos = 0
        self._done = False
        self.bytes_out = 0
        self._spare: "queue.SimpleQueue" = queue.SimpleQueue()  # chunk buffers the consumer has emptied
        self._t = threading.Thread(target=self._run, args=(producer,), daemon=True, name="mirge-read")
        self._t.start()

    def spare(self) -> Optional[bytearray]:
        """A chunk buffer (bytearray) the consumer is done with, or None: for producers that fill their own buffers."""
        try:
            return self._spare.get_nowait()
        except queue.Empty:
            return None

    def _put(self, item) -> bool:
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def _run(self, producer):
        try:
            for chunk in producer():
                if not self._put(chunk):
                    return
            self._put(None)
        except BaseException as e:  # noqa: BLE001 -- handed to the consumer
            self._put(e)

    def readinto(self, out) -> int:
        out = memoryview(out).cast("B")
        n = 0
        while n < len(out):
            if self._pos == len(self._cur):
                if self._done:
                    break
                done_with = self._cur.obj
                self._cur = memoryview(b"")
                self._pos = 0
                if isinstance(done_with, bytearray):
                    self._spare.put(done_with)  # (the view above was the last reference the consumer held)
                item = self._q.get()
                if item is None:
                    self._done = True
                    break
                if isinstance(item, BaseException):
                    self._done = True
                    raise item
                self._cur = memoryview(item)
                self._pos = 0
                continue
            k = min(len(out) - n, len(self._cur) - self._pos)
            out[n : n + k] = self._cur[self._pos : self._pos + k]
            n += k
            self._pos += k
        self.bytes_out += n
        return n

    def read(self, n: int = -1) -> bytes:
        if n is None or n < 0:
            parts = []
            while True:
                b = self.read(CHUNK)
                if not b:
                    return b"".join(parts)
                parts.append(b)
        buf = bytearray(n)
        k = self.readinto(buf)
        return bytes(buf[:k])

    def close(self):
        self._stop.set()
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass
        self._t.join(timeout=5)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class PlainReader:
    """File-like source (``readinto`` / ``read``) over an uncompressed file: ``readinto`` fills the caller's buffer --
    the pinned staging buffer of the host streamer -- with positional reads issued from the worker pool, one stripe per
    task (the read system call releases the GIL), so the bytes go from the page cache to pinned memory once, at the
    rate of several memcpy streams instead of through one reader thread and a queue of bytes objects."""

    STRIPE = 8 << 20

    def __init__(self, path: str, threads: Optional[int] = None, name: str = "", start: int = 0, stop: Optional[int] = None):
        """``start`` / ``stop``: read only that byte range of the file (a rank's share, see fastq_split_points)."""
        self.name = name or path
        self._fd = os.open(path, os.O_RDONLY)
        size = os.fstat(self._fd).st_size
        self._size = size if stop is None else max(min(int(stop), size), 0)
        self._pos = max(min(int(start), self._size), 0)
        self._threads = int(threads or default_threads())
        self.bytes_out = 0

    def _read_stripe(self, mv, offset: int) -> int:
        got = 0
        while got < len(mv):
            k = os.preadv(self._fd, [mv[got:]], offset + got)
            if k <= 0:
                break
            got += k
        return got

    def readinto(self, out) -> int:
        out = memoryview(out).cast("B")
        want = min(len(out), max(self._size - self._pos, 0))
        if want <= 0:
            return 0
        if want <= self.STRIPE or self._threads <= 1:
            got = self._read_stripe(out[:want], self._pos)
        else:
            pool = inflate_pool(self._threads)
            tasks = [(a, min(a + self.STRIPE, want)) for a in range(0, want, self.STRIPE)]
            futs = [pool.submit(self._read_stripe, out[a:b], self._pos + a) for a, b in tasks]
            got = 0
            for (a, b), f in zip(tasks, futs):
                k = f.result()
                if got == a:  # contiguous so far (a short read can only be the end of a file that shrank)
                    got += k
        self._pos += got
        self.bytes_out += got
        return got

    def read(self, n: int = -1) -> bytes:
        if n is None or n < 0:
            n = max(self._size - self._pos, 0)
        buf = bytearray(n)
        k = self.readinto(buf)
        return bytes(buf[:k])

    def close(self):
        if self._fd >= 0:
            os.close(self._fd)
            self._fd = -1

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _record_start(buf: bytes, base: int, at_eof: bool) -> Optional[int]:
    """Offset (in the file) of the first FASTQ record that starts inside ``buf`` (= the file's bytes from ``base``),
    skipping the partial line ``buf`` begins in.  A record starts with a line '@...' whose second successor is a line
    '+...' and whose first and third successors have equal lengths: a QUALITY line may begin with '@' too, but then the
    line two below it is a sequence, never '+'.  None: not decidable inside ``buf``."""
    nl = buf.find(b"\n")
    while nl >= 0:
        p = nl + 1  # start of a line
        ends = []
        q = p
        for _ in range(4):
            e = buf.find(b"\n", q)
            if e < 0:
                break
            ends.append(e)
            q = e + 1
        if len(ends) < 4:
            if at_eof and p >= len(buf):
                return base + len(buf)  # the partial line was the last one: the range ends with the file
            return None
        l0, l1, l2, l3 = (buf[a:b].rstrip(b"\r") for a, b in zip([p] + [e + 1 for e in ends[:3]], ends))
        if l0[:1] == b"@" and l2[:1] == b"+" and len(l1) == len(l3):
            return base + p
        nl = ends[0]
    return None


def fastq_split_points(path: str, parts: int, window: int = 1 << 16) -> List[int]:
    """Byte offsets [0, ..., size] that cut an uncompressed FASTQ file into ``parts`` ranges of about equal size, each
    starting at a record: what N ranks that digest ONE large sample need to read their shares themselves
    (``PlainReader(path, start=, stop=)``) instead of one process reading for all.  Compressed files cannot be cut this
    way (one DEFLATE stream has no entry points): ValueError."""
    if sniff(path) != "plain":
        raise ValueError("%s is compressed: a single compressed stream cannot be read from the middle" % path)
    size = os.path.getsize(path)
    cuts = [0]
    with open(path, "rb") as f:
        for k in range(1, int(parts)):
            target = max(size * k // int(parts), cuts[-1])
            w = int(window)
            while True:
                f.seek(target)
                buf = f.read(w)
                at_eof = target + len(buf) >= size
                # (a range that begins exactly at a record start is found too: the search starts one byte early)
                lead = 1 if target > 0 else 0
                if lead:
                    f.seek(target - 1)
                    buf = f.read(w + 1)
                pos = _record_start(buf, target - lead, at_eof) if target > 0 else 0
                if pos is not None:
                    break
                if at_eof:
                    pos = size
                    break
                w *= 4
            cuts.append(max(pos, cuts[-1]))
    cuts.append(size)
    return cuts


def open_fastq(path: str, threads: Optional[int] = None, depth: int = 4):
    """Reader for one FASTQ file: plain, gzip, BGZF, bzip2 or xz by magic number (what xopen(path, "rb") takes, digest.py:136)."""
    kind = sniff(path)
    if kind == "plain" and os.path.isfile(path):
        return PlainReader(path, threads)
    if kind == "bgzf":
        return ChunkReader(lambda: _bgzf_chunks(path, threads), depth, name=path)
    if kind == "gzip":
        big = os.path.isfile(path) and os.path.getsize(path) >= PGZ_MIN_BYTES
        if big and int(threads or default_threads()) > 1 and parallel_gzip_library() is not None:
            # the decoder works in waves of threads x PGZ_CHUNK compressed bytes and only starts the next wave when the
            # previous one has been handed over: the queue must hold a wave's output (about 4x its input for FASTQ) for
            # decoding and consumption to overlap
            wave_out = 4 * int(threads or default_threads()) * PGZ_CHUNK
            holder = []
            r = ChunkReader(lambda: _pgzip_chunks(path, threads, spare=lambda: holder[0].spare() if holder else None),
                            max(depth, -(-2 * wave_out // CHUNK)), name=path)
            holder.append(r)
            return r
        return ChunkReader(lambda: _gzip_chunks(path), depth, name=path)
    if kind == "bz2":
        import bz2

        return ChunkReader(lambda: _module_chunks(bz2.open, path), depth, name=path)
    if kind == "xz":
        import lzma

        return ChunkReader(lambda: _module_chunks(lzma.open, path), depth, name=path)
    return ChunkReader(lambda: _plain_chunks(path), depth, name=path)


def readahead_budget_bytes() -> int:
    """Host memory the look-ahead readers of a run may fill together: MIRGE_B200_READAHEAD_MB, else an eighth of
    the machine's memory, at most 16 GB."""
    env = os.environ.get("MIRGE_B200_READAHEAD_MB")
    if env:
        return max(int(env), 1) << 20
    try:
        total = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except (ValueError, OSError):
        total = 8 << 30
    return int(min(total // 8, 16 << 30))


class SampleReadahead:
    """The input files of a run, opened in order with the next ``ahead`` of them already being read and inflated.

    ``open(i)`` must be called with increasing ``i``; it returns the reader of file ``i`` (a context manager) and
    starts the readers of files ``i + 1 .. i + ahead``.  Each look-ahead reader may buffer ``budget / (ahead + 1)``
    bytes, so a cohort of single-stream gzip files is inflated by ``ahead + 1`` cores at once while the device
    digests the sample in front."""

    def __init__(self, paths: Sequence[str], ahead: Optional[int] = None, threads: Optional[int] = None,
                 budget_bytes: Optional[int] = None):
        self.paths = list(paths)
        self.threads = int(threads or default_threads())
        if ahead is None:
            ahead = min(max(self.threads - 1, 0), 8)
        self.ahead = max(0, min(int(ahead), max(len(self.paths) - 1, 0)))
        budget = int(budget_bytes if budget_bytes is not None else readahead_budget_bytes())
        self.depth = max(2, budget // ((self.ahead + 1) * CHUNK))
        self._readers: List[object] = [None] * len(self.paths)  # ChunkReader, the OSError of its open, or None
        self._next = 0

    def _start_until(self, last: int):
        while self._next <= min(last, len(self.paths) - 1):
            i = self._next
            try:
                # the readers of a run share the cores: ahead + 1 of them are at work at once
                self._readers[i] = open_fastq(self.paths[i], max(1, -(-self.threads // (self.ahead + 1))), self.depth)
            except OSError as e:  # a missing / unreadable file fails when its turn comes, not while it is looked ahead
                self._readers[i] = e
            self._next += 1

    def open(self, i: int):
        if i < 0 or i >= len(self.paths):
            raise IndexError(i)
        self._start_until(i + self.ahead)
        r = self._readers[i]
        if r is None:
            raise RuntimeError("SampleReadahead.open() must be called once per file, in order")
        self._readers[i] = None
        if isinstance(r, BaseException):
            raise r
        return r

    def close(self):
        for i, r in enumerate(self._readers):
            if isinstance(r, (ChunkReader, PlainReader)):
                r.close()
            self._readers[i] = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
