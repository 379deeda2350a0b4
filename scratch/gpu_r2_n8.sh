#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-8}
TAG=${2:-r2n8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
tail -5 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_n$N.json"))
    print(d["n_gpus"], d["value"], d["ms_per_step"], d.get("e2e"), json.dumps(d["kernels"]))
except Exception as e:
    print("bench line unreadable:", e)
PY
nvidia-smi topo -m 2>/dev/null | head -14
