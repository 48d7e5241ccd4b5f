#!/bin/bash
# round 2, call 30: evidence with the final library — GPU suite, bench lines (own + reference arm), memcheck, ncu launch list + full capture
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests.txt; cat gpurun_out/r2_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 300 gpurun_out/r2_bench_reference.json
timeout 480 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2_san_memcheck.txt 2>&1; tail -3 gpurun_out/r2_san_memcheck.txt
SICP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --pairs 8 > gpurun_out/r2_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches.csv
SICP_GRAPH=0 timeout 600 ncu --set full --clock-control none -f -o /tmp/r2_full python tools/probe_one.py > gpurun_out/r2_full.log 2>&1; tail -2 gpurun_out/r2_full.log
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null; wc -l gpurun_out/r2_full_raw.csv
