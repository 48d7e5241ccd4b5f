#!/bin/bash
# round 2, call 38: LM solve held to 176 / 160 / 144 registers per thread in a batch (other registrations' kernels co-resident on its SMs)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
STAGES=0 timeout 400 python tools/sweep.py 32 "0:37:8,2:37:8,3:37:8,4:37:8,3:37:10,4:37:10,3:44:8,3:32:8,3:37:6,0:37:8,3:37:8" 5 > gpurun_out/r2_lm_regs_sweep.txt 2>&1
cat gpurun_out/r2_lm_regs_sweep.txt
