"""usage: make_trim_cases.py [reads per case = 2000] [output file]
Writes tests/native/trim_cases.bin for tests/native/trim_check.cu: per case the trim parameters (the mirge_trim_params
bytes the product would pass), the FASTQ bytes and the windows / kept flags the C oracle computes for them.  Runs on the
build host (it uses the oracle); the GPU box only reads the file."""
import ctypes as C
import dataclasses
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mirge_b200  # noqa: E402,F401
from mirge_b200 import params as P  # noqa: E402
from oracle import coracle  # noqa: E402
from tests.test_adapter_search_host import PIPELINES  # noqa: E402
from tests.test_oracle_fuzz import random_config, random_placement_config, random_reads  # noqa: E402
from tests.util import CONFIG_DATA, CONFIGS, random_fastq  # noqa: E402


def cases(n_reads):
    # controls: forms the GPU suite has covered all round (if these fail, the harness is at fault, not the new forms)
    for name in ("default", "front_back", "linked", "wild"):
        yield "control_" + name, CONFIGS[name], random_fastq(n_reads, seed=7, n_rate=0.02, lower_rate=0.01, **CONFIG_DATA.get(name, {})), 0
    rng = np.random.default_rng(9003)
    cfg = random_config(rng)
    yield "control_fuzz_9003", cfg, random_reads(rng, cfg, n_reads), 1
    # the new forms: fixed pipelines, then the random placement configurations of the GPU test
    for i, kw in enumerate(PIPELINES):
        cfg = P.TrimConfig(**kw)
        rng = np.random.default_rng(4700 + i)
        yield "pipeline_%d" % i, cfg, random_reads(rng, cfg, n_reads), 0
    for seed in range(40):
        rng = np.random.default_rng(9500 + seed)
        cfg = random_placement_config(rng)
        yield "placement_seed_%d" % seed, cfg, random_reads(rng, cfg, n_reads), seed & 1
    # --match-read-wildcards on reads with many N: occurrences through an N in front of literal ones (match_to's shortcut)
    for seed in range(10):
        rng = np.random.default_rng(9700 + seed)
        cfg = dataclasses.replace(random_placement_config(rng), match_read_wildcards=True)
        data = bytearray(random_reads(rng, cfg, n_reads))
        lines = bytes(data).split(b"\n")
        for i in range(1, len(lines), 4):
            q = bytearray(lines[i])
            for j in np.flatnonzero(rng.random(len(q)) < 0.04):
                q[j] = ord("N")
            lines[i] = bytes(q)
        yield "readwild_seed_%d" % seed, cfg, b"\n".join(lines), seed & 1


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "trim_cases.bin")
    blobs = []
    for name, cfg, data, mode in cases(n_reads):
        cp = P.build_trim_params(cfg)
        fq = np.frombuffer(data, dtype=np.uint8)
        n, win, kept = coracle.trim(fq, cp)
        E = P.trim_slots(cp)
        pb = bytes(cp)
        blobs.append(name.encode()[:63].ljust(64, b"\0") + struct.pack("<IIQ", mode, E, len(pb)) + pb + struct.pack("<Q", len(data)) + data +
                     struct.pack("<Q", n) + win.astype("<u2").tobytes() + kept.astype(np.uint8).tobytes())
        print("%-24s mode %d  %d reads x %d slots  %s" % (name, mode, n, E, cfg.adapters))
    with open(out, "wb") as f:
        f.write(b"MRGCASE1" + struct.pack("<I", len(blobs)))
        for b in blobs:
            f.write(b)
    print("wrote %s (%d cases, %.1f MB)" % (out, len(blobs), os.path.getsize(out) / 1e6))


if __name__ == "__main__":
    main()
