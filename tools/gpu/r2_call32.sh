#!/bin/bash
# round 2, call 32: timeline of one batch (pass-by-pass path with per-stage events)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/trace_batch.py 16 8 > gpurun_out/r2_trace_batch.txt 2>&1
cat gpurun_out/r2_trace_batch.txt
