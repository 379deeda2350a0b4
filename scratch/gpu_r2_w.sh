#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "entry_points and entry]" ) > gpurun_out/pytest_multi_entry.log 2>&1
tail -15 gpurun_out/pytest_multi_entry.log
