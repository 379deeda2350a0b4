"""bench.py's output contract on the CPU arm (the only arm that runs without a GPU): exactly one JSON line on
stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "20000", "--mrna", "300"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "M reads/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_ingest_gz_leg_of_the_bench_line():
    """The host-side `ingest_gz` leg: its input really is ONE gzip member holding ONE DEFLATE stream (what makes the
    parallel decoder necessary), and the leg reports both readers on it."""
    import gzip
    import zlib

    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    from tests.util import random_fastq

    unit = random_fastq(3000, seed=2)
    blob = bench.single_stream_gzip(unit, 4)
    assert gzip.decompress(blob) == unit * 4
    d = zlib.decompressobj(31)
    assert d.decompress(blob) == unit * 4 and d.eof and d.unused_data == b""  # one member, nothing behind it
    res = bench.run_ingest_gz(np.frombuffer(unit, dtype=np.uint8), 2, target_mb=3)
    assert res["value"] > 0 and res["serial_zlib_mb_s"] > 0 and res["threads"] == 2 and res["fastq_mb"] > 2
