set -x
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
STAGES=0 timeout 400 python tools/sweep.py 16 "0:37:8,1:37:8,1:37:12,1:37:16,1:49:16,1:30:16,0:37:16" 5 2>&1 | tail -8
