"""Host-side probe (no GPU work): single-stream gzip FASTQ through ingest.open_fastq -- serial zlib reader against the
chunk-parallel decoder (csrc/pinflate.c) at several thread counts.  usage: python scratch/ingest_gzip_probe.py [MB]"""
import gzip
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mirge_b200  # noqa: E402,F401
from mirge_b200 import ingest  # noqa: E402
from tests.util import random_fastq  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 160
path = "/dev/shm/mirge_gzip_probe.fastq.gz"
t = time.perf_counter()
unit = random_fastq(40000, seed=1, L=75)  # ~7 MB of FASTQ text; repeats lie beyond the 32 KB window
reps = max(1, (mb << 20) // len(unit))
with gzip.open(path, "wb", compresslevel=4) as f:
    for _ in range(reps):
        f.write(unit)
n = reps * len(unit)
print("cores %d; file: %.0f MB of FASTQ -> %.0f MB gzip (ratio %.2f), written in %.1f s"
      % (os.cpu_count(), n / 1e6, os.path.getsize(path) / 1e6, n / os.path.getsize(path), time.perf_counter() - t), flush=True)
buf = bytearray(64 << 20)


def run(threads, parallel):
    os.environ["MIRGE_B200_PARALLEL_GZIP"] = "1" if parallel else "0"
    t0 = time.perf_counter()
    tot = 0
    with ingest.open_fastq(path, threads=threads) as r:
        while True:
            k = r.readinto(buf)
            if not k:
                break
            tot += k
    assert tot == n, (tot, n)
    return n / (time.perf_counter() - t0) / 1e6


print("serial zlib reader: %.0f MB/s" % run(1, False), flush=True)
for th in (4, 8, 16, 32, 64):
    if th <= 2 * (os.cpu_count() or 1):
        print("parallel decoder, %2d threads: %.0f MB/s" % (th, run(th, True)), flush=True)
os.unlink(path)
