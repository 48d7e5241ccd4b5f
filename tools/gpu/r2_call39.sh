#!/bin/bash
# round 2, call 39: batch default = register-capped LM CTA — GPU suite and bench lines of both arms
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests.txt; cat gpurun_out/r2_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 200 gpurun_out/r2_bench_reference.json
