#!/bin/bash
# light ncu capture (source counters + warp states + scheduler + memory workload) of selected kernels on a reduced run
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-prof}
KREGEX=${2:-'annot_search'}
READS=${3:-1000000}
timeout 420 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:"$KREGEX" -c 3 -f -o gpurun_out/prof_$TAG \
  python bench.py --reads $READS --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
