set -x
cd "$(dirname "$0")/../.."
L=$PWD/semantic-icp_b200/lib
timeout 300 python tools/sweep.py 16 "0:37:8,0:74:8,1:74:8" 4 2>&1 | tail -8
SICP_STATS_LIB=$L/libsicp_b200_stats.so timeout 200 python tools/stats.py 2>&1 | grep -E "^LM|self kNN|self k=20"
