#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2v}
( timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "sharding and 2-sharded]" ) > gpurun_out/pytest_multi_$TAG.log 2>&1
tail -3 gpurun_out/pytest_multi_$TAG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err
tail -2 gpurun_out/bench_${TAG}_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_n2.json')); print(d['n_gpus'], d['value'], d['ms_per_step'])"
