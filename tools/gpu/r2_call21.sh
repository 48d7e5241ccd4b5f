# paired LM: correctness, then throughput with pairing on / off and pair grids
set -x
cd "$(dirname "$0")/../.."
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SICP_PAIR=1 timeout 200 python - <<'PY' 2>&1 | tail -6
import numpy as np, semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
pairs = [synth.cached("kitti_pair", i, n_points=30000) for i in range(5)]
opts = sicp.default_options(sicp.ALGO_EM, cm=pairs[0]["cm"])
cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
singles = [sicp.register(sicp.ALGO_EM, s, t, opts, q["init"]) for (s, t), q in zip(cl, pairs)]
batch = sicp.register_batch(sicp.ALGO_EM, [c[0] for c in cl], [c[1] for c in cl], opts, np.stack([q["init"] for q in pairs]))
for a, b in zip(singles, batch):
    print("single", a["outer_iter"], a["lm_iters_total"], "paired", b["outer_iter"], b["lm_iters_total"], synth.pose_error(a["pose"], b["pose"]))
PY
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for pr in 0 1; do SICP_PAIR=$pr STAGES=0 timeout 300 python tools/sweep.py 16 "0:37:8,0:37:16" 5 2>&1 | tail -2; done
for g in 50 100 148; do SICP_LM_PAIR_GRID=$g STAGES=0 timeout 300 python tools/sweep.py 16 "0:37:8,0:37:16" 5 2>&1 | tail -2; done
