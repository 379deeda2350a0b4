#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2annot}
MIRGE_B200_EMPTY_CACHE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'annot_|assign_ids|drain_kernel' -c 12 -f -o gpurun_out/prof_$TAG \
  python bench.py --reads 2500000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300
python profiles/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/prof_${TAG}_summary.txt 2>&1
grep -c "Kernel Name" gpurun_out/prof_${TAG}_summary.txt
