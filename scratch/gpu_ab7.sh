#!/bin/bash
# A/B of library variants: a parity selection (PYTEST_K) on each candidate, then the short bench; the stock library last
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=mirge3.0_b200/libmirge_b200.so
cp $L /tmp/stock.so
for V in "$@" stock; do
  if [ "$V" = stock ]; then cp /tmp/stock.so $L; else cp scratch/variants/$V.so $L; fi
  touch $L
  if [ -n "$PYTEST_K" ] && [ "$V" != stock ]; then ( timeout 600 python -m pytest tests/test_gpu_annotate.py tests/test_gpu_digest.py -x -q -k "$PYTEST_K" 2>&1 | tail -2 ); fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$V.json 2> gpurun_out/ab_$V.err
done
python - "$@" stock <<'PY'
import json, sys
for n in sys.argv[1:]:
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e)
PY
