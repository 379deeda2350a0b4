#!/bin/bash
# round-end validation of the committed state: GPU parity tests, smoke, the full bench line, the ncu launch list of
# the timed region and one --set full capture of the digest kernels on a reduced run
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-v40}
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "tests rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_r1_${TAG}_full.json 2> gpurun_out/bench_r1_${TAG}_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_r1_$TAG.bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'trim|collapse_insert|tok_count|tok_index' -c 12 -f -o gpurun_out/prof_digest_$TAG \
  python bench.py --reads 2500000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_digest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_r1_${TAG}_full.json | cut -c1-400
