set -x
cd "$(dirname "$0")/../.."
L=$PWD/semantic-icp_b200/lib
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "" _nointer _hilbert; do SICP_LIB=$L/libsicp_b200$v.so python tools/probe_knn.py 2>&1 | tail -1; done
for v in "" _nointer; do SICP_LIB=$L/libsicp_b200$v.so python tools/probe_cov.py 2>&1 | tail -1; done
for v in "" _nointer; do SICP_LIB=$L/libsicp_b200$v.so STAGES=0 timeout 300 python tools/sweep.py 16 "0:37:8" 5 2>&1 | tail -2; done
