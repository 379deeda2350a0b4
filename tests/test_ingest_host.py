"""CPU tests of the host-side ingest (mirge3.0_b200/ingest.py): the readers that stand where the reference has
``xopen(FQfile, "rb")`` + ``dnaio.read_chunks`` (mirge/libs/digest.py:136-140).  Every reader must hand the device
path exactly the bytes Python's gzip module (what xopen falls back to) would: plain, gzip, multi-member gzip, zero
padding, BGZF; truncated or corrupt files must raise; the look-ahead over the samples of a run must keep the order."""
import gzip
import os
import struct
import threading
import zlib

import numpy as np
import pytest

import mirge_b200  # noqa: F401
from mirge_b200 import ingest
from tests.util import random_fastq


def write_bgzf(path, data: bytes, block=60000, eof_block=True):
    """Minimal BGZF writer (SAM spec 4.1): gzip members with a 'BC' extra field holding the block size - 1."""
    with open(path, "wb") as f:
        pieces = [data[i : i + block] for i in range(0, len(data), block)]
        if eof_block:
            pieces.append(b"")
        for p in pieces:
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = c.compress(p) + c.flush()
            bsize = 12 + 6 + len(payload) + 8
            f.write(b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1))
            f.write(payload)
            f.write(struct.pack("<II", zlib.crc32(p) & 0xFFFFFFFF, len(p)))


def read_all(reader, bufsize):
    out = bytearray()
    buf = np.zeros(bufsize, dtype=np.uint8)
    mv = memoryview(buf)  # what digest.HostStreamer hands to readinto()
    while True:
        k = reader.readinto(mv)
        if not k:
            return bytes(out)
        out += buf[:k].tobytes()


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("ingest")
    data = random_fastq(20000, seed=3)  # ~2.6 MB
    paths = {}
    p = str(d / "plain.fastq")
    open(p, "wb").write(data)
    paths["plain"] = p
    p = str(d / "single.fastq.gz")
    with gzip.open(p, "wb") as f:
        f.write(data)
    paths["gzip"] = p
    p = str(d / "members.gz")
    cut = [0, 1, 100001, len(data) // 2, len(data)]
    with open(p, "wb") as f:
        for a, b in zip(cut, cut[1:]):
            f.write(gzip.compress(data[a:b]))
        f.write(b"\x00" * 300)  # tape-style zero padding, ignored like the gzip module does
    paths["members"] = p
    p = str(d / "padded_between.gz")
    with open(p, "wb") as f:
        f.write(gzip.compress(data[:5000]) + b"\x00" * 17 + gzip.compress(data[5000:]))
    paths["padded"] = p
    p = str(d / "blocks.fastq.gz")  # BGZF under a .gz name: the format is taken from the bytes, not the suffix
    write_bgzf(p, data)
    paths["bgzf"] = p
    p = str(d / "blocks_noeof.bgz")
    write_bgzf(p, data, block=1000, eof_block=False)
    paths["bgzf_small"] = p
    p = str(d / "empty.fastq")
    open(p, "wb").close()
    paths["empty"] = p
    p = str(d / "empty.fastq.gz")
    with gzip.open(p, "wb"):
        pass
    paths["empty_gz"] = p
    return d, data, paths


def test_sniff(files):
    _, _, paths = files
    assert ingest.sniff(paths["plain"]) == "plain"
    assert ingest.sniff(paths["empty"]) == "plain"
    assert ingest.sniff(paths["gzip"]) == "gzip"
    assert ingest.sniff(paths["members"]) == "gzip"
    assert ingest.sniff(paths["bgzf"]) == "bgzf"
    assert ingest.sniff(paths["bgzf_small"]) == "bgzf"


@pytest.mark.parametrize("kind", ["plain", "gzip", "members", "padded", "bgzf", "bgzf_small"])
@pytest.mark.parametrize("bufsize", [4096, 1 << 20, 5 << 20])
def test_readers_return_the_stream(files, kind, bufsize):
    _, data, paths = files
    with ingest.open_fastq(paths[kind], threads=3, depth=2) as r:
        got = read_all(r, bufsize)
    assert got == data
    if kind != "plain":
        # what the reference's xopen / gzip would have produced
        ref = gzip.open(paths[kind], "rb").read()
        assert got == ref


def test_small_chunks_all_formats(files, monkeypatch):
    _, data, paths = files
    monkeypatch.setattr(ingest, "CHUNK", 1 << 12)
    monkeypatch.setattr(ingest, "RAW_READ", 777)
    monkeypatch.setattr(ingest, "BGZF_BATCH", 2500)
    for kind in ("plain", "gzip", "members", "padded", "bgzf", "bgzf_small"):
        with ingest.open_fastq(paths[kind], threads=4, depth=3) as r:
            assert read_all(r, 10007) == data, kind
    with ingest.open_fastq(paths["bgzf_small"], threads=2, depth=1) as r:
        head = bytearray()
        for _ in range(3000):
            b7 = bytearray(7)
            k = r.readinto(b7)
            head += b7[:k]
        assert bytes(head) == data[: len(head)] and len(head) == 21000
    with ingest.open_fastq(paths["gzip"], threads=2) as r:
        assert r.read(10) == data[:10]
        assert r.read() == data[10:]


def test_bzip2_and_xz_inputs(files, tmp_path):
    """xopen(path, "rb") also takes bzip2 and xz files (by magic number): same bytes, multi-stream files included, a
    truncated file raises."""
    import bz2
    import lzma

    _, data, _ = files
    half = len(data) // 2
    for name, mod in (("reads.fastq.bz2", bz2), ("reads.fastq.xz", lzma), ("reads_without_suffix", bz2)):
        p = str(tmp_path / name)
        with open(p, "wb") as f:
            f.write(mod.compress(data[:half]) + mod.compress(data[half:]))  # two streams
        assert ingest.sniff(p) == ("bz2" if mod is bz2 else "xz")
        with ingest.open_fastq(p, threads=2) as r:
            assert read_all(r, 1 << 20) == data
        raw = open(p, "rb").read()
        open(p, "wb").write(raw[: len(raw) // 3])
        with pytest.raises(Exception):
            with ingest.open_fastq(p, threads=2) as r:
                read_all(r, 1 << 20)


def test_fastq_split_points_and_range_readers(tmp_path):
    """Record-aligned shares of one plain FASTQ for N ranks: every cut is the start of a record although quality lines
    begin with '@' and '+' all over the file; the ranges read back to back give the file; compressed files refuse."""
    rng = np.random.default_rng(21)
    recs, starts, pos = [], [], 0
    for i in range(4000):
        L = int(rng.integers(1, 120))
        seq = "".join(rng.choice(list("ACGTN"), L))
        qual = "".join(rng.choice(list("@+IIFF#5:"), L))  # a third of the quality lines start with '@' or '+'
        if i % 7 == 0:
            qual = "@" + qual[1:]
        if i % 11 == 0:
            qual = "+" + qual[1:]
        rec = "@r%d %s\n%s\n+%s\n%s\n" % (i, "x" * int(rng.integers(0, 30)), seq, "r%d" % i if i % 5 == 0 else "", qual)
        starts.append(pos)
        pos += len(rec)
        recs.append(rec)
    data = "".join(recs).encode()
    for name, blob in (("a.fastq", data), ("crlf.fastq", data.replace(b"\n", b"\r\n")), ("noeol.fastq", data[:-1])):
        p = str(tmp_path / name)
        open(p, "wb").write(blob)
        true = set(starts)
        if b"\r" in blob:  # record starts of the CRLF file: every record is 4 bytes longer
            true = {st + 4 * i for i, st in enumerate(starts)}
        for parts in (1, 2, 3, 7, 64):
            for window in (50, 1 << 16):
                cuts = ingest.fastq_split_points(p, parts, window=window)
                assert len(cuts) == parts + 1 and cuts[0] == 0 and cuts[-1] == len(blob) and cuts == sorted(cuts)
                assert all(c in true or c == len(blob) for c in cuts), (name, parts, [c for c in cuts if c not in true])
                if parts <= 7:
                    assert all(abs((b - a) - len(blob) / parts) < 1000 for a, b in zip(cuts, cuts[1:])), (parts, cuts)
                got = b""
                for a, b in zip(cuts, cuts[1:]):
                    with ingest.PlainReader(p, threads=2, start=a, stop=b) as r:
                        got += r.read()
                assert got == blob
    tiny = str(tmp_path / "tiny.fastq")
    open(tiny, "wb").write(b"@a\nAC\n+\nII\n")
    assert ingest.fastq_split_points(tiny, 4) == [0, 11, 11, 11, 11]
    gz = str(tmp_path / "a.fastq.gz")
    with gzip.open(gz, "wb") as f:
        f.write(data)
    with pytest.raises(ValueError):
        ingest.fastq_split_points(gz, 2)


def test_empty_inputs(files):
    _, _, paths = files
    for k in ("empty", "empty_gz"):
        with ingest.open_fastq(paths[k]) as r:
            assert r.readinto(bytearray(100)) == 0
            assert r.readinto(bytearray(100)) == 0


def test_truncated_and_corrupt_inputs_raise(files):
    d, data, paths = files
    raw = open(paths["gzip"], "rb").read()
    p = str(d / "trunc.gz")
    open(p, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(EOFError):
        with ingest.open_fastq(p) as r:
            read_all(r, 1 << 20)
    with pytest.raises(EOFError):  # the gzip module agrees
        gzip.open(p, "rb").read()
    raw = open(paths["bgzf"], "rb").read()
    p = str(d / "trunc.bgz")
    open(p, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(EOFError):
        with ingest.open_fastq(p) as r:
            read_all(r, 1 << 20)
    bad = bytearray(raw)
    bad[len(raw) // 3] ^= 0x55  # flips a bit inside some block's payload
    p = str(d / "corrupt.bgz")
    open(p, "wb").write(bytes(bad))
    with pytest.raises((OSError, zlib.error)):
        with ingest.open_fastq(p) as r:
            read_all(r, 1 << 20)
    p = str(d / "garbage_after.gz")
    open(p, "wb").write(open(paths["gzip"], "rb").read() + b"not gzip")
    with pytest.raises(zlib.error):
        with ingest.open_fastq(p) as r:
            read_all(r, 1 << 20)


def test_sample_readahead_keeps_order_and_stops_its_threads(files):
    _, data, paths = files
    order = ["gzip", "plain", "bgzf", "members", "empty", "bgzf_small", "padded"]
    before = threading.active_count()
    with ingest.SampleReadahead([paths[k] for k in order], ahead=3, threads=4, budget_bytes=4 * ingest.CHUNK) as ra:
        assert ra.depth == 2  # the budget bounds what the look-ahead readers may hold
        for i, k in enumerate(order):
            with ra.open(i) as r:
                got = read_all(r, 1 << 19)
            assert got == (b"" if k == "empty" else data), k
        with pytest.raises(RuntimeError):
            ra.open(2)  # each file once, in order
    # a run that stops early must not leave readers behind
    ra = ingest.SampleReadahead([paths[k] for k in order], ahead=4, threads=4, budget_bytes=2 * ingest.CHUNK)
    r0 = ra.open(0)
    assert r0.readinto(bytearray(1000)) == 1000
    r0.close()
    ra.close()
    for t in threading.enumerate():
        if t.name == "mirge-read":
            t.join(timeout=5)
    assert sum(t.name == "mirge-read" and t.is_alive() for t in threading.enumerate()) == 0
    assert threading.active_count() <= before + ingest.default_threads() + 4  # pool workers may stay


def test_missing_file_fails_at_its_turn(files):
    d, data, paths = files
    with ingest.SampleReadahead([paths["gzip"], str(d / "nope.fastq"), paths["plain"]], ahead=2, threads=2) as ra:
        with ra.open(0) as r:
            assert read_all(r, 1 << 20) == data  # the sample in front is not affected by the look-ahead's failure
        with pytest.raises(FileNotFoundError):
            ra.open(1)
        with ra.open(2) as r:
            assert read_all(r, 1 << 20) == data


def test_readahead_defaults():
    ra = ingest.SampleReadahead(["a", "b", "c"], threads=16, budget_bytes=1 << 30)
    assert ra.ahead == 2 and ra.depth == (1 << 30) // (3 * ingest.CHUNK)
    ra = ingest.SampleReadahead(["a"], threads=16, budget_bytes=1 << 30)
    assert ra.ahead == 0
    os.environ["MIRGE_B200_READAHEAD_MB"] = "64"
    try:
        assert ingest.readahead_budget_bytes() == 64 << 20
    finally:
        del os.environ["MIRGE_B200_READAHEAD_MB"]


def test_plain_reader_parallel_positional_reads(tmp_path):
    """Uncompressed files: readinto() fills the caller's buffer with striped positional reads from the pool."""
    from mirge_b200 import ingest

    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, size=3 * ingest.PlainReader.STRIPE + 12345, dtype=np.uint8).tobytes()
    p = tmp_path / "plain.fastq"
    p.write_bytes(b"@" + data[1:])  # not a gzip magic number
    data = b"@" + data[1:]
    r = ingest.open_fastq(str(p), threads=4)
    assert isinstance(r, ingest.PlainReader)
    with r:
        out = bytearray()
        buf = bytearray(2 * ingest.PlainReader.STRIPE + 77)  # spans several stripes, not a multiple of the stripe
        while True:
            k = r.readinto(buf)
            if not k:
                break
            out += buf[:k]
        assert bytes(out) == data and r.bytes_out == len(data)
    with ingest.open_fastq(str(p), threads=1) as r1:
        assert r1.read(10) == data[:10] and r1.read() == data[10:] and r1.read(5) == b""
    empty = tmp_path / "empty.fastq"
    empty.write_bytes(b"")
    with ingest.open_fastq(str(empty)) as r0:
        assert r0.read() == b""


# ---- chunk-parallel inflate of single-stream gzip (csrc/pinflate.c -> libmirge_inflate.so) ---------------------------


def _parallel(monkeypatch, chunk_kb=64):
    if ingest.parallel_gzip_library() is None:
        pytest.skip("libmirge_inflate.so not built (run __graft_entry__.build())")
    monkeypatch.setattr(ingest, "PGZ_MIN_BYTES", 0)
    monkeypatch.setattr(ingest, "PGZ_CHUNK", chunk_kb << 10)


@pytest.mark.parametrize("kind", ["gzip", "members", "padded", "empty_gz"])
@pytest.mark.parametrize("threads,chunk_kb", [(2, 64), (5, 64), (8, 256)])
def test_parallel_gzip_reader_returns_the_stream(files, monkeypatch, kind, threads, chunk_kb):
    """The native chunk-parallel decoder gives exactly what Python's gzip module (the reference's xopen fallback) gives,
    for single streams, concatenated members (incl. 1-byte and empty ones) and zero padding."""
    d, data, paths = files
    _parallel(monkeypatch, chunk_kb)
    r = ingest.open_fastq(paths[kind], threads=threads)
    with r:
        got = read_all(r, 777_777)
    assert got == gzip.open(paths[kind], "rb").read()


def test_parallel_gzip_reader_on_every_block_type(tmp_path, monkeypatch):
    """Stored, fixed-Huffman and dynamic blocks, sync / full flush points, tiny blocks (memLevel 1), runs with distance 1,
    incompressible bytes, a gzip header with a file name: many speculative chunks per file, every one chained and
    CRC-checked against the trailer."""
    _parallel(monkeypatch, 64)
    rng = np.random.default_rng(5)
    fq = random_fastq(60000, seed=8)

    def gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8):
        c = zlib.compressobj(level, zlib.DEFLATED, 31, memlevel, strategy)
        return c.compress(data) + c.flush()

    def flushed(data, step=300_000):
        c = zlib.compressobj(6, zlib.DEFLATED, 31)
        out = []
        for i in range(0, len(data), step):
            out.append(c.compress(data[i : i + step]))
            out.append(c.flush(zlib.Z_FULL_FLUSH if (i // step) % 2 else zlib.Z_SYNC_FLUSH))
        return b"".join(out) + c.flush()

    import io

    named = io.BytesIO()
    with gzip.GzipFile(filename="sample_R1.fastq", mode="wb", fileobj=named, mtime=0) as f:
        f.write(fq[:2_000_000])
    cases = {
        "level1": gz(fq, 1), "level9": gz(fq, 9), "fixed": gz(fq[:3_000_000], strategy=zlib.Z_FIXED),
        "huffman_only": gz(fq[:3_000_000], strategy=zlib.Z_HUFFMAN_ONLY), "rle": gz(fq[:3_000_000], strategy=zlib.Z_RLE),
        "stored": gz(fq[:3_000_000], 0), "small_blocks": gz(fq[:4_000_000], memlevel=1), "zeros": gz(bytes(20_000_000)),
        "random": gz(rng.integers(0, 256, 2_000_000, dtype=np.uint8).tobytes()), "repeats": gz(b"ACGTACGTTTGACCA\n" * 500_000, 9),
        "flushes": flushed(fq[:4_000_000]), "named": named.getvalue(),
        "mixed_members": gz(fq[:2_000_000]) + gz(fq[2_000_000:2_000_001]) + gz(b"") + gz(fq[2_000_001:5_000_000], 1) + bytes(99),
    }
    for name, comp in cases.items():
        p = tmp_path / (name + ".gz")
        p.write_bytes(comp)
        with ingest.open_fastq(str(p), threads=6) as r:
            got = read_all(r, 1 << 20)
        assert got == gzip.decompress(comp), name


def test_parallel_gzip_reader_errors(files, monkeypatch):
    d, data, paths = files
    _parallel(monkeypatch, 64)
    raw = open(paths["gzip"], "rb").read()

    def reading(blob, name):
        p = str(d / name)
        open(p, "wb").write(blob)
        with ingest.open_fastq(p, threads=4) as r:
            return read_all(r, 1 << 20)

    with pytest.raises(EOFError):
        reading(raw[: len(raw) // 2], "p_trunc.gz")
    with pytest.raises(EOFError):
        reading(raw[:-3], "p_trunc_trailer.gz")
    bad = bytearray(raw)
    bad[len(raw) // 2] ^= 0x10
    with pytest.raises((OSError, zlib.error)):  # invalid code, or the CRC at the end
        reading(bytes(bad), "p_flip.gz")
    bad = bytearray(raw)
    bad[-6] ^= 1
    with pytest.raises(OSError):
        reading(bytes(bad), "p_crc.gz")
    bad = bytearray(raw)
    bad[-1] ^= 1
    with pytest.raises(OSError):
        reading(bytes(bad), "p_isize.gz")
    with pytest.raises((OSError, zlib.error)):
        reading(raw + b"not gzip", "p_garbage.gz")


def test_parallel_gzip_is_only_used_for_large_files(files, monkeypatch):
    d, data, paths = files
    if ingest.parallel_gzip_library() is None:
        pytest.skip("libmirge_inflate.so not built")
    calls = []
    real = ingest._pgzip_chunks
    monkeypatch.setattr(ingest, "_pgzip_chunks", lambda *a, **k: (calls.append(a), real(*a, **k))[1])
    with ingest.open_fastq(paths["gzip"], threads=4) as r:  # 1 MB: below PGZ_MIN_BYTES
        read_all(r, 1 << 20)
    assert not calls
    monkeypatch.setattr(ingest, "PGZ_MIN_BYTES", 0)
    with ingest.open_fastq(paths["gzip"], threads=1) as r:  # one thread: nothing to parallelise
        read_all(r, 1 << 20)
    assert not calls
    with ingest.open_fastq(paths["gzip"], threads=4) as r:
        assert read_all(r, 1 << 20) == data
    assert len(calls) == 1
    monkeypatch.setenv("MIRGE_B200_PARALLEL_GZIP", "0")
    with ingest.open_fastq(paths["gzip"], threads=4) as r:
        assert read_all(r, 1 << 20) == data
    assert len(calls) == 1


@pytest.mark.parametrize("seed", range(8))
def test_parallel_gzip_reader_fuzz(tmp_path, monkeypatch, seed):
    """Random data (text, low entropy, periodic, incompressible, mixtures), random members / levels / strategies / window
    memory / flush points, random thread counts and cut sizes: the parallel decoder returns what zlib returns."""
    _parallel(monkeypatch, 64)
    rng = np.random.default_rng(700 + seed)

    def rand_data(n, depth=0):
        kind = int(rng.integers(0, 5 if depth == 0 else 4))
        if kind == 0:
            return random_fastq(n // 150 + 1, seed=int(rng.integers(1 << 30)))[:n]
        if kind == 1:
            return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        if kind == 2:
            return rng.integers(0, 4, n, dtype=np.uint8).tobytes()
        if kind == 3:
            unit = rng.integers(65, 91, int(rng.integers(1, 5000)), dtype=np.uint8).tobytes()
            return (unit * (n // len(unit) + 1))[:n]
        parts, tot = [], 0
        while tot < n:
            k = int(rng.integers(1, 150_000))
            parts.append(rand_data(k, 1))
            tot += k
        return b"".join(parts)[:n]

    for case in range(3):
        data = rand_data(int(rng.integers(1, 2_500_000)))
        cuts = sorted({0, len(data)} | {int(x) for x in rng.integers(0, len(data) + 1, int(rng.integers(0, 3)))})
        blob = []
        for a, b in zip(cuts, cuts[1:]):
            c = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, 31, int(rng.integers(1, 10)),
                                 int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])))
            p = a
            while p < b:
                k = int(rng.integers(1, 300_000))
                blob.append(c.compress(data[p : min(p + k, b)]))
                p += k
                if rng.random() < 0.2:
                    blob.append(c.flush(int(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))))
            blob.append(c.flush())
            if rng.random() < 0.3:
                blob.append(bytes(int(rng.integers(1, 40))))
        comp = b"".join(blob)
        assert gzip.decompress(comp) == data
        p = tmp_path / ("f%d_%d.gz" % (seed, case))
        p.write_bytes(comp)
        monkeypatch.setattr(ingest, "PGZ_CHUNK", int(rng.choice([64, 128, 512])) << 10)
        with ingest.open_fastq(str(p), threads=int(rng.integers(2, 9))) as r:
            got = read_all(r, int(rng.integers(1, 1 << 21)))
        assert got == data, (seed, case)


def test_inflate_library_is_built_and_exports_its_entry_points():
    """libmirge_inflate.so (csrc/pinflate.c) is built by __graft_entry__.build() and exports what include/mirge_inflate.h declares."""
    import ctypes

    if not os.path.exists(ingest.PGZ_LIB):
        pytest.skip("libmirge_inflate.so not built (run __graft_entry__.build())")
    import re

    lib = ctypes.CDLL(ingest.PGZ_LIB)
    header = open(os.path.join(os.path.dirname(ingest.__file__), "..", "include", "mirge_inflate.h")).read()
    declared = set(re.findall(r"\b(pgz_[a-z]+)\s*\(", header))
    assert declared == {"pgz_open", "pgz_read", "pgz_error", "pgz_stats", "pgz_times", "pgz_close", "pgz_crc"}
    for name in declared:  # every entry point include/mirge_inflate.h declares
        assert hasattr(lib, name), name


def test_parallel_gzip_reader_closed_early_and_without_prefetch(files, monkeypatch):
    """The decoder's producer thread works one wave ahead of the reader: closing a reader that has only taken a few bytes
    (or nothing) must stop and join it; MIRGE_B200_PGZ_PREFETCH=0 (waves decoded on demand) returns the same bytes."""
    _, data, paths = files
    _parallel(monkeypatch, 64)
    for take in (0, 1, 70000):
        r = ingest.open_fastq(paths["gzip"], threads=4)
        assert r.read(take) == data[:take]
        r.close()
    before = threading.active_count()
    for _ in range(5):
        with ingest.open_fastq(paths["members"], threads=3) as r:
            assert r.read(1000) == data[:1000]
    assert threading.active_count() <= before
    monkeypatch.setenv("MIRGE_B200_PGZ_PREFETCH", "0")
    for kind in ("gzip", "members", "padded"):
        with ingest.open_fastq(paths[kind], threads=4) as r:
            assert read_all(r, 1 << 20) == data


def test_crc_paths_equal_zlib():
    """pgz_crc (carry-less multiplication for the bulk, slice-by-8 for the rest and on CPUs without PCLMULQDQ) against
    zlib.crc32: every length around the 16- and 64-byte steps, odd alignments, running CRCs over pieces."""
    import ctypes
    import subprocess
    import sys

    if not os.path.exists(ingest.PGZ_LIB):
        pytest.skip("libmirge_inflate.so not built (run __graft_entry__.build())")
    code = r"""
import ctypes, sys, zlib, numpy as np
lib = ctypes.CDLL(sys.argv[1])
lib.pgz_crc.restype = ctypes.c_uint32
lib.pgz_crc.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64]
rng = np.random.default_rng(1)
buf = rng.integers(0, 256, 70000, dtype=np.uint8)
base = buf.ctypes.data
n = 0
for off in (0, 1, 3, 8, 13):
    for ln in list(range(0, 300)) + [1000, 4095, 4096, 4097, 32768, 65519]:
        assert lib.pgz_crc(0, base + off, ln) == zlib.crc32(buf[off:off + ln].tobytes()), (off, ln)
        n += 1
crc_a = crc_b = 0
pos = 0
while pos < buf.size:
    k = int(rng.integers(0, 500))
    crc_a = lib.pgz_crc(crc_a, base + pos, min(k, buf.size - pos))
    crc_b = zlib.crc32(buf[pos:pos + k].tobytes(), crc_b)
    pos += k
assert crc_a == crc_b
print("ok", n)
"""
    for no_clmul in ("0", "1"):  # the switch is read once per process
        env = dict(os.environ, MIRGE_B200_PGZ_NO_CLMUL=no_clmul)
        out = subprocess.run([sys.executable, "-c", code, ingest.PGZ_LIB], env=env, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and out.stdout.startswith("ok"), (no_clmul, out.stdout, out.stderr)
