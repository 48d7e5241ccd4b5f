#!/bin/bash
# round 2, call 40: covariance kernel held to 80 registers (6 CTAs per SM) beside the register-capped LM
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=$PWD/semantic-icp_b200/lib
{
for lib in "" $L/libsicp_b200_cov6.so "" $L/libsicp_b200_cov6.so; do
  echo "== ${lib:-default}"
  SICP_LIB=$lib STAGES=0 timeout 300 python tools/sweep.py 32 "3:37:8" 6 2>&1 | tail -1
done
} > gpurun_out/r2_cov6.txt 2>&1
cat gpurun_out/r2_cov6.txt
