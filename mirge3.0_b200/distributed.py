"""Multi-GPU collapse: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The path shards naturally: FASTQ chunks / samples are split across ranks, every
rank collapses locally, and exactly one exchange makes each rank the owner of a disjoint slice of the
unique sequences: records (count, packed key) are partitioned by hash64(centre sequence) mod world
(kernel: mirge_partition_plan), packed per destination (mirge_partition_pack), moved with a single
all-to-all, and merged into the owner's table (mirge_collapse_merge).  With UMIs the centre (UMI flanks
removed) is hashed, so both UMI levels are owner-local.  Libraries are replicated; annotation then runs on
the owner's slice with no further communication.  The reference has no counterpart: its only
parallelism is a fork pool whose results are merged by one Python loop (digest.py:139-163)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.distributed as dist


def shard_ranges(total: int, world: int):
    """Contiguous [lo, hi) unit ranges (chunks, samples, reads) per rank; sizes differ by at most 1."""
    base, rem = divmod(int(total), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def exchange_records(rec: torch.Tensor, sizes: torch.Tensor, send_recs: torch.Tensor, group=None,
                     send_words: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-to-all of variable-size records.

    rec       int32 words of all records, grouped by destination rank (ascending)
    sizes     int32 [n] words of each record, same order
    send_recs int64 [world] number of records for each destination
    send_words optional int64 [world] words for each destination (computed from sizes when absent)
    Returns (recv words, recv sizes) with records grouped by source rank.  Device-agnostic plumbing:
    NCCL on GPUs, gloo in the CPU tests."""
    world = dist.get_world_size(group)
    dev = rec.device
    send_recs = send_recs.to(torch.int64)
    if send_words is None:  # words per destination from the record sizes
        bounds = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        bounds[1:] = torch.cumsum(send_recs.to(dev), 0)
        csum = torch.zeros(sizes.numel() + 1, dtype=torch.int64, device=dev)
        csum[1:] = torch.cumsum(sizes.to(torch.int64), 0)
        send_words = csum[bounds[1:]] - csum[bounds[:-1]]
    send_words = send_words.to(torch.int64).to(dev)
    meta_out = torch.stack([send_recs.to(dev), send_words]).t().contiguous()  # [world, 2]
    meta_in = torch.empty_like(meta_out)
    dist.all_to_all_single(meta_in, meta_out, group=group)
    meta_in_h = meta_in.cpu()
    recv_recs = [int(x) for x in meta_in_h[:, 0]]
    recv_words = [int(x) for x in meta_in_h[:, 1]]
    s_recs = [int(x) for x in send_recs.cpu()]
    s_words = [int(x) for x in send_words.cpu()]
    r_sizes = torch.empty(sum(recv_recs), dtype=sizes.dtype, device=dev)
    dist.all_to_all_single(r_sizes, sizes.contiguous(), output_split_sizes=recv_recs, input_split_sizes=s_recs, group=group)
    r_rec = torch.empty(sum(recv_words), dtype=rec.dtype, device=dev)
    dist.all_to_all_single(r_rec, rec.contiguous(), output_split_sizes=recv_words, input_split_sizes=s_words, group=group)
    return r_rec, r_sizes


def exchange_and_merge(dev, local_table, ids: torch.Tensor, cnt: torch.Tensor, owner_table, world: int, umi=(0, 0), group=None):
    """Partition the drained (key id, count) pairs of ``local_table`` by owner, exchange, merge into
    ``owner_table`` (GPU path; every kernel through the C ABI): plan (owner and record size of every pair) ->
    totals per owner -> scatter into per-owner regions of the send buffer (cursors, no sort) -> all-to-all ->
    merge."""
    from .device import _ptr

    lib = dev.lib
    n = int(ids.numel())
    dest = dev.empty(n, torch.int32)
    words = dev.empty(n, torch.int32)
    totals = dev.zeros(2 * world, torch.int64)
    with dev.timed("xchg_plan"):
        if n:
            dev.check(lib.mirge_partition_plan(dev.ctx, C.byref(local_table.struct), _ptr(ids), n, int(umi[0]), int(umi[1]), world,
                                               _ptr(dest), _ptr(words), dev.stream()))
        dev.check(lib.mirge_partition_totals(dev.ctx, _ptr(dest), _ptr(words), n, world, _ptr(totals), dev.stream()))
        dev.launches += 2
        tot = totals.cpu()
    send_words, send_recs = tot[:world].clone(), tot[world:].clone()
    total = int(send_words.sum())
    if total >= (1 << 31):
        raise RuntimeError("exchange larger than 2^31 words per rank; lower the batch size")
    base = ((torch.cumsum(send_recs, 0) - send_recs) << 32) | (torch.cumsum(send_words, 0) - send_words)
    rec = dev.empty(total, torch.int32)
    sizes = dev.empty(n, torch.int32)
    with dev.timed("xchg_pack"):
        cursors = base.to(dev.tdev)
        if n:
            dev.check(lib.mirge_partition_scatter(dev.ctx, C.byref(local_table.struct), _ptr(ids), _ptr(cnt), _ptr(dest), _ptr(words), n,
                                                  world, _ptr(cursors), _ptr(rec), _ptr(sizes), dev.stream()))
            dev.launches += 1
    with dev.timed("xchg_a2a"):
        r_rec, r_sizes = exchange_records(rec[:total], sizes[:n], send_recs, group, send_words=send_words)
    m = int(r_sizes.numel())
    if m == 0:
        return 0
    with dev.timed("xchg_merge"):
        r_off = (torch.cumsum(r_sizes.to(torch.int64), 0) - r_sizes.to(torch.int64)).to(torch.int32)
        owner_table.check()
        owner_table.reserve(m, int(r_rec.numel()))
        deferred = dev.empty(m, torch.int32)
        dev.check(lib.mirge_collapse_merge(dev.ctx, C.byref(owner_table.struct), _ptr(r_rec), _ptr(r_off), m, _ptr(deferred), dev.stream()))
        dev.launches += 3
        owner_table.check()
    return m
