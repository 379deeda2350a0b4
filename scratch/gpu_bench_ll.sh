#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-x}
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}.json"))
    print(d["value"], d["ms_per_step"], d.get("e2e"), json.dumps(d["kernels"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
bash scratch/gpu_launchlist.sh $TAG | head -24
