set -x
cd "$(dirname "$0")/../.."
for pr in 0 1; do SICP_PAIR=$pr timeout 300 python tools/sweep.py 16 "0:37:8" 5 2>&1 | grep -E "batch of|variant"; done
SICP_PAIR=1 SICP_LM_PAIR_GRID=148 timeout 300 python tools/sweep.py 16 "0:37:8" 5 2>&1 | grep -E "batch of|variant"
SICP_PAIR=1 SICP_LM_PAIR_GRID=74 SICP_CONCURRENT=16 timeout 300 python tools/sweep.py 16 "0:37:16" 5 2>&1 | grep -E "variant"
