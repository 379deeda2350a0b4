#!/bin/bash
# bisect: several library variants on one box; parity tests + short bench each
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=mirge3.0_b200/libmirge_b200.so
for V in "$@"; do
  cp scratch/variants/$V.so $L; touch $L
  if [ "$V" = "$1" ]; then T="tests/test_gpu_digest.py tests/test_gpu_annotate.py tests/test_gpu_report.py tests/test_gpu_fullsize.py"; else T="tests/test_gpu_digest.py tests/test_gpu_annotate.py"; fi
  ( time timeout 600 python -m pytest $T -x -q ) > gpurun_out/ab_tests_$V.log 2>&1
  echo "tests rc=$?" >> gpurun_out/ab_tests_$V.log
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$V.json 2> gpurun_out/ab_$V.err
done
python - "$@" <<'PY'
import json, sys
for n in sys.argv[1:]:
    print(n, open("gpurun_out/ab_tests_%s.log" % n).read().strip().splitlines()[-1])
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e)
PY
