#!/bin/bash
# round 2, call 24: the paired-solve test alone (full traceback), then the round's bench lines (own arm + reference arm)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k paired 2>&1 | tail -60 > gpurun_out/r2_paired_test.txt
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 3000 gpurun_out/r2_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
tail -c 1500 gpurun_out/r2_bench_reference.json
