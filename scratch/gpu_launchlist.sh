#!/bin/bash
# ncu launch list (durations only) of the timed region of a bench run; TAG names the outputs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-ll}
READS=${2:-50000000}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --reads $READS \
  > gpurun_out/bench_ncu_$TAG.json 2> gpurun_out/bench_ncu_$TAG.err
python profiles/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_$TAG.txt 2>&1
cat gpurun_out/launches_$TAG.txt
# per-launch order for the annotate kernels (round by round)
python - <<PY
import csv
rows = list(csv.reader(l for l in open("gpurun_out/launches_$TAG.csv") if l.startswith('"')))
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[1:]:
    n = r[ki].split("(")[0]
    if "annot" in n or "collapse" in n or "assign" in n:
        print(n[:40], r[vi], r[ui])
PY
