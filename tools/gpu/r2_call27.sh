#!/bin/bash
# round 2, call 27: f64-compare insertion chain — parity tests, kernel rates and batch throughput against the integer-compare build
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests_f64cmp.txt; cat gpurun_out/r2_gpu_tests_f64cmp.txt
L=$PWD/semantic-icp_b200/lib
{
for rep in 1 2; do
  timeout 200 python tools/probe_knn.py
  SICP_LIB=$L/libsicp_b200_icmp.so timeout 200 python tools/probe_knn.py
done
STAGES=1 timeout 300 python tools/sweep.py 32 "0:37:8" 5 2>&1 | tail -5
SICP_LIB=$L/libsicp_b200_icmp.so STAGES=1 timeout 300 python tools/sweep.py 32 "0:37:8" 5 2>&1 | tail -5
SICP_LIB=$L/libsicp_b200_p1fma.so timeout 200 python tools/probe_knn.py
SICP_LIB=$L/libsicp_b200_p1fma.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k knn 2>&1 | tail -3
} > gpurun_out/r2_call27_ab.txt 2>&1
cat gpurun_out/r2_call27_ab.txt
