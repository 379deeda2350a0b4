#!/bin/bash
# A/B of library variants on one box: parity tests on the candidate, then the short bench on candidate and base.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=mirge3.0_b200/libmirge_b200.so
CAND=${1:-all2}
cp scratch/variants/$CAND.so $L; touch $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab_gpu.txt
( time timeout 900 python -m pytest tests/test_gpu_digest.py tests/test_gpu_annotate.py tests/test_gpu_report.py tests/test_gpu_fullsize.py -x -q --durations=10 ) > gpurun_out/ab_tests_$CAND.log 2>&1
echo "tests rc=$?" >> gpurun_out/ab_tests_$CAND.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$CAND.json 2> gpurun_out/ab_$CAND.err
cp scratch/variants/base.so $L; touch $L
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
cp scratch/variants/$CAND.so $L; touch $L
tail -5 gpurun_out/ab_tests_$CAND.log
python - <<PY
import json
for n in ("$CAND","base"):
    try:
        d=json.loads(open("gpurun_out/ab_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
