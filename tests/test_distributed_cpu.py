"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): sharding + the variable-size
record all-to-all (mirge_b200.distributed.exchange_records).  Records are built by the oracle's
collapse here; on GPUs the same plumbing moves records packed by mirge_partition_pack."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mirge_b200  # noqa: F401
from mirge_b200 import distributed as MD
from mirge_b200 import params as P
from oracle import coracle
from tests.util import CONFIGS, random_fastq


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def owner_of(key: str, world: int) -> int:
    return zlib.crc32(key.encode()) % world


def pack(records):
    """[(key, count)] -> (int32 words, int32 sizes): record = [count][len][ascii bytes, 4 per word]"""
    words, sizes = [], []
    for k, c in records:
        b = k.encode() + b"\0" * (-len(k) % 4)
        w = [c, len(k)] + list(np.frombuffer(b, dtype=np.int32))
        words.extend(int(x) for x in w)
        sizes.append(len(w))
    return torch.tensor(words, dtype=torch.int32), torch.tensor(sizes, dtype=torch.int32)


def unpack(words, sizes):
    out, o = [], 0
    w = words.numpy()
    for s in sizes.tolist():
        c, ln = int(w[o]), int(w[o + 1])
        out.append((w[o + 2 : o + s].tobytes()[:ln].decode(), c))
        o += s
    return out


def worker(rank, world, port, data, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = CONFIGS["release"]
        cp = P.build_trim_params(cfg)
        # shard whole records across ranks
        recs = data.split(b"\n@")
        recs = [recs[0]] + [b"@" + r for r in recs[1:]]
        lo, hi = MD.shard_ranges(len(recs), world)[rank]
        shard = b"\n".join(recs[lo:hi])
        if not shard.endswith(b"\n"):
            shard += b"\n"
        _, tab = coracle.digest_collapse(np.frombuffer(shard, dtype=np.uint8), cp)
        local = tab.to_dict()
        by_dest = sorted(local.items(), key=lambda kv: (owner_of(kv[0], world), kv[0]))
        words, sizes = pack(by_dest)
        send = torch.zeros(world, dtype=torch.int64)
        for k, _ in by_dest:
            send[owner_of(k, world)] += 1
        r_words, r_sizes = MD.exchange_records(words, sizes, send)
        merged = {}
        for k, c in unpack(r_words, r_sizes):
            assert owner_of(k, world) == rank
            merged[k] = merged.get(k, 0) + c
        q.put((rank, merged))
    finally:
        dist.destroy_process_group()


def test_shard_ranges():
    assert MD.shard_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert MD.shard_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert MD.shard_ranges(0, 2) == [(0, 0), (0, 0)]


def test_hash_partitioned_exchange_world2():
    world = 2
    data = random_fastq(1500, seed=9, pool=80)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, data, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cp = P.build_trim_params(CONFIGS["release"])
    _, tab = coracle.digest_collapse(np.frombuffer(data, dtype=np.uint8), cp)
    exp = tab.to_dict()
    # every unique sequence is owned by exactly one rank and the merged counts equal the global collapse
    assert set(results[0]).isdisjoint(results[1])
    union = {**results[0], **results[1]}
    assert union == exp
    assert len(results[0]) > 0 and len(results[1]) > 0
