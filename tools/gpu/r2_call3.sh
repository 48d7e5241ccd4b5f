# round 2, call 3: LM shape sweep with graph exec reuse, curve-order A/B in a batch, traversal statistics of both curve orders
set -x
cd "$(dirname "$0")/../.."
L=$PWD/semantic-icp_b200/lib
timeout 500 python tools/sweep.py 16 "0:37:8,0:37:12,0:37:16,0:49:12,0:74:12,1:74:12,1:148:12,1:148:16,1:296:12,2:74:12,2:148:12,2:296:12,3:148:12,3:296:12,4:74:12,4:148:12,0:37:8:0" 5 2>&1 | tail -18
SICP_LIB=$L/libsicp_b200_morton.so timeout 200 python tools/sweep.py 16 "0:37:8,0:37:12" 5 2>&1 | tail -3
SICP_STATS_LIB=$L/libsicp_b200_stats.so timeout 200 python tools/stats.py 2>&1 | tail -14
SICP_STATS_LIB=$L/libsicp_b200_stats_morton.so timeout 200 python tools/stats.py 2>&1 | tail -14
