#!/bin/bash
# round 2, call 36: final library — GPU suite, bench lines of both arms, ncu launch list + full capture (LM and search kernels changed since call 30)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests.txt; cat gpurun_out/r2_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 300 gpurun_out/r2_bench_reference.json
SICP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --pairs 8 > gpurun_out/r2_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches.csv
SICP_GRAPH=0 timeout 600 ncu --set full --clock-control none -f -o /tmp/r2_full python tools/probe_one.py > gpurun_out/r2_full.log 2>&1; tail -2 gpurun_out/r2_full.log
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null; wc -l gpurun_out/r2_full_raw.csv
SICP_GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:lm_kernel -s 12 -c 4 -f -o /tmp/r2_lmb python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --pairs 8 > gpurun_out/r2_lmb.log 2>&1
ncu -i /tmp/r2_lmb.ncu-rep --page raw --csv > gpurun_out/r2_lm_batch_raw.csv 2>/dev/null; wc -l gpurun_out/r2_lm_batch_raw.csv
