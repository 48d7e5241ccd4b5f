set -x
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
python tools/probe_cov.py 2>&1 | tail -1
timeout 300 python tools/sweep.py 16 "0:37:8" 4 2>&1 | tail -5
