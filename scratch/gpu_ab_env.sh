#!/bin/bash
# A/B of environment-controlled variants inside one call: tests once, then a short resident bench per variant
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-ab}
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "tests rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
for v in 1 4 16 64; do
  MIRGE_B200_HK_ITEMS=$v timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ab_${TAG}_hk$v.json 2> gpurun_out/ab_${TAG}_hk$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${TAG}_hk$v.json"))
    print("hk=$v", d["value"], d["ms_per_step"], {k: x["ms_per_step"] for k, x in d["kernels"].items()})
except Exception as e:
    print("hk=$v failed", e)
PY
done
