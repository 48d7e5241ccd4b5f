#!/bin/bash
# round 2, call 37 (8 GPUs): the driver's own multi-GPU launch with the final library
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; cut -c1-400 gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err
