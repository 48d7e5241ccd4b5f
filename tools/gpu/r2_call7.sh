set -x
cd "$(dirname "$0")/../.."
L=$PWD/semantic-icp_b200/lib
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/probe_knn.py 2>&1 | tail -1
python tools/probe_cov.py 2>&1 | tail -1
timeout 300 python tools/sweep.py 16 "0:37:8" 4 2>&1 | tail -2
