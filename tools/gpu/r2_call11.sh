set -x
cd "$(dirname "$0")/../.."
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 2 --pairs 32 > gpurun_out/r2_c11_bench2.json 2> gpurun_out/r2_c11_bench2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_c11_bench2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','ms_per_step_by_rank','records_gathered')}, d['e2e']['value'], d['config']['outer_passes_hist'])"; tail -3 gpurun_out/r2_c11_bench2.err
timeout 900 $TR tools/run_configs.py c4 --pairs 96 2>&1 | tail -2
timeout 900 $TR tools/run_configs.py c5 --pairs 4 --inits 256 2>&1 | tail -2
