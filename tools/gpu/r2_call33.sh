#!/bin/bash
# round 2, call 33 (2 GPUs): the driver's own multi-GPU launch of both arms with the final library
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 1500 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2_reference.json 2> gpurun_out/r2_bench_n2_reference.err; tail -c 400 gpurun_out/r2_bench_n2_reference.json
