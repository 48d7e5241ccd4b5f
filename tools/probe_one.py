"""One KITTI-shaped EM registration (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
p = synth.kitti_pair(0, n_points=n)
src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
r = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
print("outer", r["outer_iter"], "lm", r["lm_iters_total"], synth.pose_error(r["pose"], p["T_gt"]))
