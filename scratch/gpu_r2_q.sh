#!/bin/bash
# full GPU suite on one device (after the compat switch), log kept
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2q}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -6 gpurun_out/pytest_gpu_$TAG.log
