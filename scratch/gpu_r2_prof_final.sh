#!/bin/bash
# round-2 evidence set: launch list of the timed region at full size, one --set full capture of every hot kernel on a reduced
# run, SASS of the bulk-copy fed kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2final}
bash scratch/gpu_launchlist.sh $TAG > gpurun_out/ll_$TAG.log 2>&1
tail -25 gpurun_out/launches_$TAG.txt
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'trim_kernel|trim_dp|collapse_list_kernel|assign_ids|annot_|tok_count|tok_index|drain_kernel' -c 40 -f -o gpurun_out/prof_$TAG \
  python bench.py --reads 2500000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
python profiles/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/prof_${TAG}_summary.txt 2>&1
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/ncu_$TAG.log") if l.startswith("{")][-1])
print("reads", d["config"]["reads_per_gpu"], "unique", d["config"]["unique_sequences"])
open("gpurun_out/prof_${TAG}_counts.json", "w").write(json.dumps({"reads": d["config"]["reads_per_gpu"], "unique": d["config"]["unique_sequences"]}))
PY
# the bulk-copy fed fused kernel (opt-in): one capture of it on the same reduced run
MIRGE_B200_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'digest_tiles' -c 2 -f -o gpurun_out/prof_${TAG}_fused \
  python bench.py --reads 2500000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${TAG}_fused.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_${TAG}_fused.ncu-rep > gpurun_out/prof_${TAG}_fused_summary.txt 2>&1
python profiles/ncu_lines_id.py gpurun_out/prof_${TAG}_fused.ncu-rep 0 40 >> gpurun_out/prof_${TAG}_fused_summary.txt 2>&1
python profiles/ncu_lines_id.py gpurun_out/prof_${TAG}_fused.ncu-rep 0 30 inst >> gpurun_out/prof_${TAG}_fused_summary.txt 2>&1
tail -5 gpurun_out/prof_${TAG}_fused_summary.txt
