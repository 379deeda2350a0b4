"""Run under torchrun (one rank per GPU): every rank digests its shard of one synthetic sample, the
unique sequences are hash-partitioned to their owners with the NCCL all-to-all, owners annotate their
slice; rank 0 gathers everything and compares with the single-process oracle on the whole sample.
Used by tests/test_gpu_multi.py; exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mirge_b200  # noqa: E402,F401
from mirge_b200 import device as D  # noqa: E402
from mirge_b200 import distributed as MD  # noqa: E402
from mirge_b200 import libraries as LB  # noqa: E402
from mirge_b200 import manifoldAlign as MA  # noqa: E402
from mirge_b200 import synth  # noqa: E402


def check_entry_points(rank, world, dev, umi):
    """baking_sharded + bwtAlign_sharded (one process per GPU) against the single-process entry points."""
    import gzip
    import tempfile

    from mirge_b200 import digest as DG
    from tests.test_gpu_annotate import make_libs, write_lib_dir
    from tests.util import ILL, make_args, random_fastq

    box = [tempfile.mkdtemp(prefix="mirge_mg_") if rank == 0 else None]
    dist.broadcast_object_list(box, 0)
    tmp = box[0]
    names = ["s%d" % i for i in range(5)]
    files = [os.path.join(tmp, n + (".fastq.gz" if i == 1 else ".fastq")) for i, n in enumerate(names)]
    if rank == 0:
        rng = np.random.default_rng(3)
        libs = make_libs(rng, scale=0.4)
        write_lib_dir(os.path.join(tmp, "lib"), libs)
        mir = libs["mirna"][1]
        for i, f in enumerate(files):
            r2 = np.random.default_rng(50 + i)
            recs = []
            for k in range(2500 + 300 * i):
                ins = mir[int(r2.zipf(1.3)) % len(mir)]
                pre = "".join(r2.choice(list("ACGT"), 4)) if umi else ""
                suf = "".join(r2.choice(list("ACGT"), 4)) if umi else ""
                s_ = (pre + ins + suf + ILL + "ACGTACGTACGTAC")[:60]
                recs.append("@r%d.%d\n%s\n+\n%s\n" % (i, k, s_, "I" * len(s_)))
            data = "".join(recs).encode()
            if f.endswith(".gz"):
                with gzip.open(f, "wb") as fh:
                    fh.write(data)
            else:
                open(f, "wb").write(data)
    dist.barrier()
    args = make_args(libraries_path=os.path.join(tmp, "lib"), spikeIn=True, uniq_mol_ids="4,4" if umi else None, umiDedup=bool(umi),
                     tcf_out=True)
    df, src, trc, tru = MD.baking_sharded(args, files, names, tmp, device=dev, batch_bytes=200_000)
    out = MD.bwtAlign_sharded(args, df, tmp, "miRBase")
    dist.barrier()

    def tcf_files():
        """{sample: [(count, sequence)]} of the <sample>.trim.collapse.fa files (-tcf): counts must descend; ties may come in
        any order (they follow the table's slot order)"""
        got = {}
        for n in names:
            lines = open(os.path.join(tmp, n + ".trim.collapse.fa")).read().split("\n")
            recs = [(int(h.rsplit("_", 1)[1]), s_) for h, s_ in zip(lines[0::2], lines[1::2]) if h]
            assert all(recs[i][0] >= recs[i + 1][0] for i in range(len(recs) - 1))
            assert [h for h in lines[0::2] if h] == [">seq%d_%d" % (i + 1, c) for i, (c, _) in enumerate(recs)]
            got[n] = sorted(recs)
        return got

    ok = True
    if rank == 0:
        tcf_sharded = tcf_files()  # written by the ranks that digested the samples
        df1, src1, trc1, tru1 = DG.baking(args, files, names, tmp, device=dev, batch_bytes=200_000)
        out1 = MA.bwtAlign(args, df1, tmp, "miRBase", device=dev)
        ok = out.equals(out1) and list(out.columns) == list(out1.columns) and (src, trc, tru) == (src1, trc1, tru1)
        ok = ok and tcf_sharded == tcf_files()
        print("sharded entry points: world=%d samples=%d rows=%d annotated=%d umi=%s ok=%s"
              % (world, len(names), len(out), int((out.annotFlag == 1).sum()), bool(umi), ok))
        if not ok:
            print(src, src1, trc, trc1, tru, tru1, out.shape, out1.shape)
    return ok


def exchange_unique(dev, eng, shard, local, world, umi, n_expected):
    """per-batch exchange of unique sequences on the worker thread / side stream behind the next batch's trim (two local
    tables, distributed.ExchangeWorker)"""
    local_t = D.CollapseTable(dev, min_keys=1 << 12)
    local_b = D.CollapseTable(dev, min_keys=1 << 12)
    nb = torch.tensor([D.DigestEngine.max_batches(shard.numel(), 3 << 20)], device=dev.tdev)
    dist.all_reduce(nb, op=dist.ReduceOp.MAX)
    worker = MD.ExchangeWorker(local, world, owner_min_keys=1 << 12, umi=umi)
    n = eng.digest_device_exchange(shard, (local_t, local_b), worker, batch_bytes=3 << 20, n_batches=int(nb.item()))
    assert n == n_expected
    return worker.finish()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = D.Device(local)
    if len(sys.argv) > 1 and sys.argv[1] in ("entry", "entry_umi"):
        ok = check_entry_points(rank, world, dev, sys.argv[1] == "entry_umi")
        flag = torch.tensor([1 if ok else 0], device=dev.tdev)
        dist.broadcast(flag, 0)
        dist.destroy_process_group()
        sys.exit(0 if int(flag.item()) == 1 else 1)
    mode = sys.argv[2] if len(sys.argv) > 2 else "unique"  # unique | sharded | sharded_sync
    sharded = mode.startswith("sharded")
    cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n_reads = 60_000 if sharded else 120_000
    libs = synth.make_libraries(scale=0.05, mrna_count=100)
    lset = LB.LibrarySet.from_fasta_dict(dev, libs.fasta_dict())
    cfg = synth.trim_config_for(cfg_id)
    eng = D.DigestEngine(dev, cfg)
    umi = cfg.umi() or (0, 0)
    # the same sample on every rank (seeded), each rank takes its contiguous share of the records
    fq = synth.ReadGenerator(libs, synth.CONFIGS[cfg_id], "cpu").fastq(n_reads).numpy()
    nl = np.flatnonzero(fq == 10)
    lo, hi = MD.shard_ranges(n_reads, world)[rank]
    b0 = 0 if lo == 0 else int(nl[4 * lo - 1]) + 1
    b1 = int(nl[4 * hi - 1]) + 1
    shard = torch.from_numpy(fq[b0:b1].copy()).to(dev.tdev)
    if sharded:
        # sharding before the collapse: unequal shards (rank 0 has more batches than the others), small batches
        cut = [0] + [int(n_reads * (0.55 + 0.45 * r / (world - 1))) for r in range(world - 1)] + [n_reads] if world > 1 else [0, n_reads]
        lo, hi = cut[rank], cut[rank + 1]
        b0 = 0 if lo == 0 else int(nl[4 * lo - 1]) + 1
        shard = torch.from_numpy(fq[b0 : int(nl[4 * hi - 1]) + 1].copy()).to(dev.tdev)
        owner_t = D.CollapseTable(dev, min_keys=1 << 12)
        sc = MD.ShardedCollapse(eng, world, overlap=(mode == "sharded"))
        n = sc.digest_device(shard, owner_t, batch_bytes=2 << 20)
        assert n == hi - lo, (n, lo, hi)
        assert sc.rounds >= 2
    else:
        owner_t = exchange_unique(dev, eng, shard, local, world, umi, hi - lo)
    ids, cnt = owner_t.drain()
    keys = owner_t.export_keys()
    annot, hit = MA.annotate_keys(dev, lset, MA.KeySet.from_table(owner_t), True)
    mine = {keys[i].decode(): (int(c), int(a), int(h)) for i, c, a, h in
            zip(ids.cpu().tolist(), cnt.cpu().tolist(), annot.cpu()[ids.cpu().long()].tolist(),
                hit.cpu().numpy().view(np.uint64)[ids.cpu().numpy()].tolist())}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        from oracle import coracle

        union = {}
        for g in gathered:
            assert not (set(g) & set(union)), "a sequence is owned by two ranks"
            union.update(g)
        _, tab = coracle.digest_collapse(fq, dev.trim_params, nthreads=8)
        exp = tab.to_dict()
        ok = {k: v[0] for k, v in union.items()} == exp
        seqs = sorted(exp)
        blob = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        a_o = np.full(len(seqs), 0xFF, dtype=np.uint8)
        h_o = np.full(len(seqs), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
        pols = LB.round_policies()
        for rnd in range(10):
            L = libs.libs[LB.ROUND_LIBS[rnd]]
            coracle.annotate_round(blob, off, lut[L.codes], L.off.astype(np.uint32), pols[rnd], a_o, h_o, nthreads=8)
        for i, s in enumerate(seqs):
            if union[s][1] != int(a_o[i]) or union[s][2] != int(h_o[i]):
                ok = False
                print("annotation mismatch", s, union[s], int(a_o[i]), int(h_o[i]))
                break
        sizes = [len(g) for g in gathered]
        print("multi-GPU check: world=%d reads=%d uniques=%d per-rank=%s annotated=%d ok=%s"
              % (world, n_reads, len(exp), sizes, int((a_o != 0xFF).sum()), ok))
    flag = torch.tensor([1 if ok else 0], device=dev.tdev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
