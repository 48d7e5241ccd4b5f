# round 2, call 4: where does the batch time go (host issue vs wait), curve order on the lone kernels, new bench line
set -x
cd "$(dirname "$0")/../.."
L=$PWD/semantic-icp_b200/lib
SICP_HOSTSTAT=1 timeout 300 python tools/sweep.py 16 "0:37:8,0:37:8:0,0:37:16" 3 2>&1 | tail -16
python tools/probe_knn.py 2>&1 | tail -1
SICP_LIB=$L/libsicp_b200_morton.so python tools/probe_knn.py 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 2 --pairs 32 > gpurun_out/r2_c4_bench.json 2> gpurun_out/r2_c4_bench.err; tail -c 3000 gpurun_out/r2_c4_bench.json; tail -5 gpurun_out/r2_c4_bench.err
