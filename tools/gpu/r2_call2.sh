# round 2, call 2: correctness of the new paths, then A/B sweeps (graph loop, LM shapes, curve order)
set -x
cd "$(dirname "$0")/../.."
L=semantic-icp_b200/lib
# smoke under a short timeout first: a graph loop that never ends must not eat the box
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
SICP_GRAPH=0 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/sweep.py 16 "0:37:8:0,0:37:8:1:0,0:37:8:1:1,0:37:12:1:0,0:37:16:1:0" 5 2>&1 | tail -8
timeout 400 python tools/sweep.py 16 "1:74:8,1:74:12,1:74:16,1:148:12,2:74:12,2:148:12,2:37:12,3:74:12,3:148:12,4:37:12,4:74:12,0:74:12,0:49:12" 4 2>&1 | tail -14
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python tools/sweep.py 16 "0:37:8,0:37:16,1:74:16" 4 2>&1 | tail -4
python tools/probe_cov.py 2>&1 | tail -1
SICP_LIB=$PWD/$L/libsicp_b200_morton.so python tools/probe_cov.py 2>&1 | tail -1
SICP_LIB=$PWD/$L/libsicp_b200_morton.so timeout 200 python tools/sweep.py 16 "0:37:8" 4 2>&1 | tail -2
timeout 200 python tools/trace_batch.py 16 8 > gpurun_out/r2_c2_trace.log 2>&1; tail -12 gpurun_out/r2_c2_trace.log
