"""Host-side ingest throughput (no GPU): single-stream gzip, block-parallel BGZF, look-ahead over a cohort of gzip files.
usage: python scratch/ingest_probe.py [reads_per_file] [files]"""
import gzip, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mirge_b200  # noqa
from mirge_b200 import ingest
from tests.util import random_fastq
from tests.test_ingest_host import write_bgzf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 6
d = tempfile.mkdtemp(prefix="ingest_probe_")
data = random_fastq(n, seed=1)
mb = len(data) / 1e6
gz = os.path.join(d, "a.fastq.gz")
with gzip.open(gz, "wb", compresslevel=4) as f:
    f.write(data)
bg = os.path.join(d, "a.bgz")
write_bgzf(bg, data)
buf = np.zeros(64 << 20, dtype=np.uint8)
def drain(r):
    t = 0
    while True:
        k = r.readinto(memoryview(buf))
        if not k:
            return t
        t += k
t0 = time.perf_counter(); k = len(gzip.open(gz, "rb").read()); t1 = time.perf_counter()
print("gzip module, one thread          : %7.1f MB/s" % (mb / (t1 - t0)))
for name, path in (("ingest gzip (1 worker)", gz), ("ingest BGZF (%d workers)" % ingest.default_threads(), bg)):
    t0 = time.perf_counter()
    with ingest.open_fastq(path) as r:
        assert drain(r) == len(data)
    t1 = time.perf_counter()
    print("%-33s: %7.1f MB/s" % (name, mb / (t1 - t0)))
paths = []
for i in range(nf):
    p = os.path.join(d, "s%d.fastq.gz" % i)
    os.link(gz, p)
    paths.append(p)
for ahead in (0, nf - 1):
    t0 = time.perf_counter()
    with ingest.SampleReadahead(paths, ahead=ahead) as ra:
        for i in range(nf):
            with ra.open(i) as r:
                assert drain(r) == len(data)
    t1 = time.perf_counter()
    print("%d gzip files, look-ahead %d       : %7.1f MB/s" % (nf, ahead, nf * mb / (t1 - t0)))
print("cores:", os.cpu_count(), " file: %.0f MB inflated" % mb)
