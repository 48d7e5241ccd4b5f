set -x
cd "$(dirname "$0")/../.."
for s in 8 4 0; do SICP_LM_CTL_SHARE=$s python tools/probe_cov.py 2>&1 | tail -1; done
for s in 8 4 0; do SICP_LM_CTL_SHARE=$s STAGES=0 timeout 300 python tools/sweep.py 16 "0:37:12" 5 2>&1 | tail -1; done
