#!/bin/bash
# sharding before the collapse: loop-back parity on one device, NCCL parity on two, bench at N ranks
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r2s}
( timeout 600 python -m pytest tests/test_gpu_digest.py -x -q -k "sharding_before or exchange_world2" ) > gpurun_out/pytest_shard_$TAG.log 2>&1
tail -15 gpurun_out/pytest_shard_$TAG.log
grep -q "failed\|error" gpurun_out/pytest_shard_$TAG.log && exit 1
( timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ${MULTI_K:+-k "$MULTI_K"} ) > gpurun_out/pytest_multi_$TAG.log 2>&1
tail -15 gpurun_out/pytest_multi_$TAG.log
grep -q "failed\|error" gpurun_out/pytest_multi_$TAG.log && exit 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
tail -5 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_n$N.json"))
    print(d["n_gpus"], d["value"], d["ms_per_step"], d.get("e2e"), json.dumps(d["kernels"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
