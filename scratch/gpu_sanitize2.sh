#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_digest.py -k "default-0 or nextseq-0" -x -q > gpurun_out/blocking.log 2>&1
grep -n "Error\|error\|passed\|failed\|mirge_" gpurun_out/blocking.log | head -20
timeout 600 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_digest.py -k "nextseq-0" -x -q > gpurun_out/racecheck_fused.log 2>&1
grep -B1 -A6 "Race\|hazard\|ERROR SUMMARY\|passed\|failed" gpurun_out/racecheck_fused.log | grep -v "Host Frame" | head -60
