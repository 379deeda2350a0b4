#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2z}
timeout 900 python bench.py --dropin > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print(d["value"], d["ms_per_step"], "e2e", d.get("e2e"), "dropin", d.get("dropin"))
PY
bash scratch/gpu_r2_prof_final.sh $TAG
