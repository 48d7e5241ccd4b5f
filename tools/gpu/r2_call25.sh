#!/bin/bash
# round 2, call 25: full GPU suite; lone wall; LM launch priority A/B; E-step register-cap variants
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests.txt; cat gpurun_out/r2_gpu_tests.txt
PROBE_WALL=1 timeout 120 python tools/probe_one.py 2>&1 | tail -5 > gpurun_out/r2_lone_wall.txt; cat gpurun_out/r2_lone_wall.txt
{
for prio in 0 -1 -5; do
  echo "== SICP_LM_PRIO=$prio"
  SICP_LM_PRIO=$prio STAGES=0 timeout 300 python tools/sweep.py 32 "0:37:8,0:37:12" 5 2>&1 | tail -3
done
for lib in em5 em6; do
  echo "== lib $lib"
  SICP_LIB=$PWD/semantic-icp_b200/lib/libsicp_b200_$lib.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --pairs 16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernels']['estep'])"
done
echo "== lib default"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --pairs 16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernels']['estep'])"
} > gpurun_out/r2_call25_ab.txt 2>&1
cat gpurun_out/r2_call25_ab.txt
