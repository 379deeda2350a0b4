#!/bin/bash
# the whole GPU suite, then bench lines for the configurations given (default: 2 with the drop-in leg, then 1 3 4 5)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-cfg}
shift
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "tests rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
for c in ${@:-2 1 3 4 5}; do
  extra=""
  if [ "$c" = "2" ]; then extra="--dropin"; fi
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 $extra > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err
  echo "config $c rc=$?"; tail -2 gpurun_out/bench_${TAG}_c$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_c$c.json"))
    print("C$c", d["value"], d["ms_per_step"], "e2e", d.get("e2e"), "dropin", d.get("dropin"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print("   ", {k: v["ms_per_step"] for k, v in d["kernels"].items()}, d["config"]["unique_sequences"], d["config"]["annotated_sequences"])
except Exception as e:
    print("C$c bench line unreadable:", e)
PY
done
