set -x
cd "$(dirname "$0")/../.."
STAGES=0 timeout 600 python tools/sweep.py 32 "0:37:8,0:37:16,0:24:16,0:18:16,0:18:24,0:24:24,0:12:24,0:12:32,0:18:32" 4 2>&1 | tail -10
