#!/bin/bash
# round 2: GPU suite + bench line of the current tree (TAG names the outputs)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2b}
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "tests rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -5 gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/bench_${TAG}.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}.json"))
    print(d["value"], d["ms_per_step"], d.get("e2e"), json.dumps(d["kernels"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
