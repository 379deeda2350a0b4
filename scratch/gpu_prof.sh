#!/bin/bash
# one --set full capture of the hot kernels on a reduced run (2.5 M reads); TAG names the outputs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-prof}
KREGEX=${2:-'trim_kernel|trim_dp|collapse_list_kernel|assign_ids|annotate_tile|digest_tile'}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -c 14 -f -o gpurun_out/prof_$TAG \
  python bench.py --reads 2500000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
