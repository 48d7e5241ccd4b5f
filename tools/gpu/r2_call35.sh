#!/bin/bash
# round 2, call 35: sweep carried as 2b, aa^2 - be^2 in one FMA, third-order rcp/rsqrt steps — parity, lone stage times and batch throughput against the previous build
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests_sweep.txt; cat gpurun_out/r2_gpu_tests_logtab.txt
L=$PWD/semantic-icp_b200/lib
{
for lib in "" $L/libsicp_b200_prev.so "" $L/libsicp_b200_prev.so; do
  echo "== lib ${lib:-default}"
  SICP_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --pairs 32 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['kernels']['lm'], d['config']['lm_iters_mean'])"
done
} > gpurun_out/r2_call35_ab.txt 2>&1
cat gpurun_out/r2_call35_ab.txt
