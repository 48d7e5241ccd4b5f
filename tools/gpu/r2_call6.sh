set -x
cd "$(dirname "$0")/../.."
timeout 200 python tools/diag_gate.py 2>&1 | tail -34
