"""BASELINE.json's full single-GPU size (C2: 50 M reads, L = 75, -a illumina -nxt 20 -q 20, HEAD counting) through
size-independent properties: record / emission conservation, independence of the batch partition, count doubling
on a second pass (no new keys), oracle parity on slices taken from three places of the stream, and annotation that
does not depend on the processing order.  Runs in well under a minute on a B200."""
import ctypes as C

import numpy as np
import pytest

from oracle import coracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

N_READS = 50_000_000


def table_signature(dev, table, ids, cnt):
    """(number of keys, sum of counts, order-independent checksum of (key hash, count)) of a drained table."""
    from mirge_b200.device import _ptr

    n = int(ids.numel())
    dest = dev.empty(n, torch.int32)
    words = dev.empty(n, torch.int32)
    dev.check(dev.lib.mirge_partition_plan(dev.ctx, C.byref(table.struct), _ptr(ids), n, 0, 0, 2147483629, _ptr(dest), _ptr(words),
                                           dev.stream()))
    h = dest.to(torch.int64)
    return n, int(cnt.to(torch.int64).sum().item()), int(((h * cnt.to(torch.int64)) % 1000000007).sum().item()), \
        int((words.to(torch.int64) * 7 + h).sum().item())


def test_c2_full_size_properties():
    from mirge_b200 import device as D
    from mirge_b200 import libraries as LB
    from mirge_b200 import manifoldAlign as MA
    from mirge_b200 import synth
    from tests.test_gpu_digest import gpu_windows

    dev = D.Device(0)
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~60 GB of free device memory")
    libs = synth.make_libraries(mrna_count=2000)
    lset = LB.LibrarySet.from_fasta_dict(dev, libs.fasta_dict())
    eng = D.DigestEngine(dev, synth.trim_config_for(2, "head"))
    fq = synth.ReadGenerator(libs, synth.CONFIGS[2], dev.tdev).fastq(N_READS)
    nbytes = int(fq.numel())

    # pass A: 2 GB batches
    ta = D.CollapseTable(dev, min_keys=1 << 24)
    for k in eng.stats:
        eng.stats[k] = 0
    assert eng.digest_device(fq, ta, 2048 << 20) == N_READS
    emitted = eng.stats["emitted"]
    assert eng.stats["records"] == N_READS and eng.stats["bytes"] == nbytes
    ids, cnt = ta.drain()
    sig_a = table_signature(dev, ta, ids, cnt)
    assert sig_a[1] == emitted and N_READS <= emitted <= 3 * N_READS  # every kept slot is counted exactly once
    assert sig_a[0] == ta.n_keys

    # pass B: a different partition of the stream into batches gives the same table
    tb = D.CollapseTable(dev, min_keys=1 << 24)
    assert eng.digest_device(fq, tb, 777 << 20) == N_READS
    ids_b, cnt_b = tb.drain()
    assert table_signature(dev, tb, ids_b, cnt_b) == sig_a
    del tb, ids_b, cnt_b

    # the same reads again into table A: no new key, every count doubles (the drain above zeroed the counts, so
    # digest twice and compare with 2x)
    eng.digest_device(fq, ta, 2048 << 20)
    eng.digest_device(fq, ta, 1500 << 20)
    ids2, cnt2 = ta.drain()
    sig2 = table_signature(dev, ta, ids2, cnt2)
    assert sig2[0] == sig_a[0] and sig2[1] == 2 * sig_a[1] and ta.n_keys == sig_a[0]

    # annotation of the 39 M sequences: independent of the processing order, and every annotated key has a hit
    keys = MA.KeySet.from_table(ta)
    a1, h1 = MA.annotate_keys(dev, lset, keys, False, split=True)   # several pre-pass chunks at this size
    a2, h2 = MA.annotate_keys(dev, lset, keys, False, split=False)  # the fused single launch
    assert torch.equal(a1, a2) and torch.equal(h1, h2)
    assert torch.equal(a1 != 0xFF, h1 != -1) and int((a1 != 0xFF).sum()) > 1_000_000
    del ta, keys, a1, a2, h1, h2, ids, cnt, ids2, cnt2

    # oracle parity of the trim windows on three 100 k-read slices (start, middle, end of the stream)
    nl = None
    for frac in (0.0, 0.5, 0.995):
        lo = int(nbytes * frac)
        chunk = fq[lo : lo + 24_000_000].cpu().numpy()
        if lo:  # align to a record start: '\n@SYN' after a quality line can only be a header here (qualities never hold '\n')
            pos = np.flatnonzero((chunk[:-1] == 10) & (chunk[1:] == ord("@")))
            # a quality line may start with '@': take the first candidate whose line after next starts with '+'
            nl = np.flatnonzero(chunk == 10)
            start = None
            for p in pos[:8].tolist():
                k = np.searchsorted(nl, p)
                if k + 2 < nl.size and chunk[nl[k + 2] + 1] == ord("+") and chunk[p + 1 : p + 5].tobytes() == b"@SYN":
                    start = p + 1
                    break
            assert start is not None
            chunk = chunk[start:]
        nl = np.flatnonzero(chunk == 10)
        n_take = min(100_000, nl.size // 4)
        chunk = np.ascontiguousarray(chunk[: nl[4 * n_take - 1] + 1])
        n, win_o, kept_o = coracle.trim(chunk, dev.trim_params, nthreads=8)
        buf = torch.from_numpy(chunk).to(dev.tdev)
        br = eng.trim_batch(buf, buf.numel(), True)
        win_g, kept_g = gpu_windows(eng, br)
        assert br.n_records == n == n_take
        assert np.array_equal(kept_g, kept_o) and np.array_equal(win_g, win_o), frac
