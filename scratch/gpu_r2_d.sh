#!/bin/bash
cd "$GRAFT_REPO_ROOT"
bash scratch/gpu_r2_b.sh $1
bash scratch/gpu_prof.sh $1
