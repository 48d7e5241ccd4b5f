set -x
cd "$(dirname "$0")/../.."
L=$PWD/semantic-icp_b200/lib
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/sweep.py 16 "0:37:8" 4 2>&1 | tail -5
SICP_LIB=$L/libsicp_b200_nopool.so timeout 300 python tools/sweep.py 16 "0:37:8" 4 2>&1 | tail -5
SICP_EAGER_BUILD=1 timeout 300 python tools/sweep.py 16 "0:37:8" 4 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 2 --pairs 32 --no-cpu-baseline --no-extra > gpurun_out/r2_c10_bench.json 2> gpurun_out/r2_c10_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_c10_bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['kernels'], d['knn_queries_per_s'])"; tail -3 gpurun_out/r2_c10_bench.err
SICP_EAGER_BUILD=1 timeout 600 python bench.py --steps 5 --warmup 2 --pairs 32 --no-cpu-baseline --no-extra 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('EAGER', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])"
