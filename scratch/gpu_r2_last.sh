#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_last.log 2>&1
tail -4 gpurun_out/smoke_last.log
timeout 300 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
tail -2 gpurun_out/bench_last.err
python -c "
import json; d=json.load(open('gpurun_out/bench_last.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['roofline'], d['gpu_launches'], d['clocks'])"
