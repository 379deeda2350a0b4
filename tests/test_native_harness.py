"""The native (no Python, no torch) parity check of the trim path, tests/native/trim_check.cu, stays buildable and keeps
reading what tests/native/make_trim_cases.py writes: built here with nvcc, run in its --dry mode (cases parsed and
size-checked against the header's structures, the library's entry points resolved, no device needed).  On a GPU box the
same binary without --dry is the check (profiles/r2_native_trim_check_b200.txt)."""
import os
import shutil
import subprocess
import sys

import pytest

import mirge_b200
from mirge_b200 import abi

ROOT = mirge_b200.REPO_ROOT
NATIVE = os.path.join(ROOT, "tests", "native")


def test_native_trim_check_builds_and_reads_its_cases(tmp_path):
    nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)
    if nvcc is None:
        pytest.skip("nvcc not found")
    if not os.path.exists(abi.LIB_PATH):
        pytest.skip("libmirge_b200.so not built (run __graft_entry__.build())")
    exe = str(tmp_path / "trim_check")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", exe,
                           os.path.join(NATIVE, "trim_check.cu"), "-ldl"])
    cases = str(tmp_path / "cases.bin")
    out = subprocess.run([sys.executable, os.path.join(NATIVE, "make_trim_cases.py"), "40", cases], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    n_cases = sum(1 for ln in out.stdout.splitlines() if " mode " in ln)
    assert n_cases >= 60
    dry = subprocess.run([exe, abi.LIB_PATH, cases, "--dry"], capture_output=True, text=True, timeout=120)
    assert dry.returncode == 0, dry.stdout[-2000:] + dry.stderr[-2000:]
    lines = dry.stdout.splitlines()
    assert sum(1 for ln in lines if ln.startswith("case ")) == n_cases and lines[-1].startswith("ALL PASS: %d cases" % n_cases)
    for form in ("control_default", "pipeline_0", "placement_seed_39", "readwild_seed_9"):
        assert any(form in ln for ln in lines), form
