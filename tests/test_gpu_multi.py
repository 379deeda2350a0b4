"""Two-GPU parity of the sharded path (NCCL all-to-all exchange + owner-side annotation) against the
single-process oracle.  Needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cfg_id", [1, 3])
def test_two_gpu_exchange_matches_oracle(cfg_id):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + cfg_id), os.path.join(ROOT, "tests", "multi_gpu_check.py"), str(cfg_id)]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "ok=True" in p.stdout


@pytest.mark.parametrize("mode", ["entry", "entry_umi"])
def test_two_gpu_entry_points_match_single_process(mode):
    """baking_sharded / bwtAlign_sharded (samples dealt to the ranks, hash-partitioned collapse, owner-side annotation,
    gather to rank 0) give the DataFrame and counters of the single-process baking / bwtAlign."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_check.py"), mode]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "ok=True" in p.stdout


@pytest.mark.parametrize("cfg_id,mode", [(1, "sharded"), (2, "sharded"), (2, "sharded_sync")])
def test_two_gpu_sharding_before_collapse_matches_oracle(cfg_id, mode):
    """distributed.ShardedCollapse over NCCL: unequal shards of one sample on two GPUs, the keys of every batch sent to
    their owners before the collapse (all-to-all on a side stream / in line), owner-side annotation; the union of the
    owners' tables and annotations equals the single-process oracle on the whole sample."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29620 + cfg_id), os.path.join(ROOT, "tests", "multi_gpu_check.py"), str(cfg_id), mode]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "ok=True" in p.stdout
