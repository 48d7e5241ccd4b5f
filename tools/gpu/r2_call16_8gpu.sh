# round 2: the 8-GPU run — bench line at N = 8 (distinct shards), C4 (1,000-pair sequence) and C5 (64 x 4,096 sweep) at full size
set -x
cd "$(dirname "$0")/../.."
nvidia-smi -L | wc -l; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 2 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','ms_per_step_by_rank','records_gathered')}, d['e2e']['value'], d['config']['outer_passes_hist'])"; tail -3 gpurun_out/r2_bench_n8.err
timeout 1200 $TR --master-port 29522 tools/run_configs.py c4 --pairs 1000 --chunk 25 > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err; tail -c 1500 gpurun_out/r2_c4_n8.json; tail -3 gpurun_out/r2_c4_n8.err
timeout 1500 $TR --master-port 29523 tools/run_configs.py c5 --pairs 64 --inits 4096 > gpurun_out/r2_c5_n8.json 2> gpurun_out/r2_c5_n8.err; tail -c 1800 gpurun_out/r2_c5_n8.json; tail -3 gpurun_out/r2_c5_n8.err
