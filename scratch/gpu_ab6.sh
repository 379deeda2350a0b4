#!/bin/bash
# parity of the fused kernel on a variant library (its seven cases + the synthetic workloads), then the bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=mirge3.0_b200/libmirge_b200.so
cp $L /tmp/stock.so
for V in "$@"; do
  cp scratch/variants/$V.so $L; touch $L
  ( MIRGE_B200_FUSED=1 timeout 600 python -m pytest tests/test_gpu_digest.py -x -q -k "fused or synthetic_workloads" 2>&1 | tail -2 )
  MIRGE_B200_FUSED=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$V.json 2> gpurun_out/ab_$V.err
done
cp /tmp/stock.so $L; touch $L
python - "$@" <<'PY'
import json, sys
for n in sys.argv[1:]:
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], {k: v["ms_per_step"] for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "bench failed", e)
PY
