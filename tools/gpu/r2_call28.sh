#!/bin/bash
# round 2, call 28: insertion-chain microbenchmark (tools/ubench/topk_insert.cu, built in-tree before the call)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 120 tools/ubench/topk_insert > gpurun_out/r2_ubench_topk.txt 2>&1
cat gpurun_out/r2_ubench_topk.txt
