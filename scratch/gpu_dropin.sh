#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-dropin}
( timeout 900 python -m pytest tests/test_gpu_report.py tests/test_gpu_annotate.py -q -x 2>&1 | tail -3 )
timeout 900 python bench.py --config 2 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --dropin > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print(d["value"], d["ms_per_step"], "dropin", d.get("dropin"))
PY
