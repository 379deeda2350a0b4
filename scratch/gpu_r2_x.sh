#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 130 python -m pytest tests/test_gpu_annotate.py -x -q -k "full_size_libraries" ) > gpurun_out/pytest_fullannot.log 2>&1
tail -12 gpurun_out/pytest_fullannot.log
