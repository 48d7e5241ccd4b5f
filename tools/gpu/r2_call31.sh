#!/bin/bash
# round 2, call 31: next-entry prefetch in the packet search — parity, kernel rates, lone wall and batch throughput against the build without it
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests_pf.txt; cat gpurun_out/r2_gpu_tests_pf.txt
L=$PWD/semantic-icp_b200/lib
{
for rep in 1 2; do
  timeout 200 python tools/probe_knn.py
  SICP_LIB=$L/libsicp_b200_nopf.so timeout 200 python tools/probe_knn.py
done
STAGES=1 timeout 300 python tools/sweep.py 32 "0:37:8" 5 2>&1 | tail -5
SICP_LIB=$L/libsicp_b200_nopf.so STAGES=1 timeout 300 python tools/sweep.py 32 "0:37:8" 5 2>&1 | tail -5
SAN_STAGES=search timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -8
} > gpurun_out/r2_call31_ab.txt 2>&1
cat gpurun_out/r2_call31_ab.txt
