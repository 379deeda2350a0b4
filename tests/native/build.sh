#!/bin/bash
# builds tests/native/trim_check (the native parity check of the trim path; see trim_check.cu)
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o trim_check trim_check.cu -ldl
