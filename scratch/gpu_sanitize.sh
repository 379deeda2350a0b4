#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke_v40.log 2>&1
echo "memcheck smoke rc=$?" >> gpurun_out/memcheck_smoke_v40.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_digest.py -x -q -k "batching or format or long_records or hot_key or multi_sample or host_streamer" > gpurun_out/memcheck_digest_v40.log 2>&1
echo "memcheck digest rc=$?" >> gpurun_out/memcheck_digest_v40.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke_v40.log 2>&1
echo "racecheck smoke rc=$?" >> gpurun_out/racecheck_smoke_v40.log
tail -4 gpurun_out/memcheck_smoke_v40.log; tail -4 gpurun_out/memcheck_digest_v40.log; tail -4 gpurun_out/racecheck_smoke_v40.log
