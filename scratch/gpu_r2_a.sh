#!/bin/bash
# round 2, call A: the whole GPU suite with the fuzz un-gated and cfg 4/5 added, plus a fresh baseline bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2a}
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "tests rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -5 gpurun_out/pytest_gpu_$TAG.log; cut -c1-600 gpurun_out/bench_${TAG}.json
