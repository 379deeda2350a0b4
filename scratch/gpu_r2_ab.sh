#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2ab}
( time timeout 1500 python -m pytest tests/test_gpu_zz_fuzz.py tests/test_gpu_digest.py -m gpu -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -6 gpurun_out/pytest_gpu_$TAG.log
grep -n "^FAILED\|^E  " gpurun_out/pytest_gpu_$TAG.log | head -20
