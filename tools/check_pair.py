"""GPU vs oracle on one pair of the soak sequence (a hard / degenerate pair: many passes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
from oracle import oracle as O
sicp, synth = pkg.sicp, pkg.synth
i = int(sys.argv[1])
frames, poses, cm = synth.kitti_sequence(i + 2, n_points=60_000, n_rings=64, n_az=940)
(sx, sl), (tx, tl) = frames[i + 1], frames[i]
ident = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
r = sicp.register(sicp.ALGO_EM, sicp.Cloud(sx, sl), sicp.Cloud(tx, tl), sicp.default_options(sicp.ALGO_EM, cm=cm), ident)
t0 = time.time(); ref = O.align_em(sx, sl, tx, tl, cm, ident); dt = time.time() - t0
print("pair", i, "gpu passes", r["outer_iter"], "oracle passes", ref["outer_iter"], "(oracle %.1f s)" % dt)
print("pose diff gpu vs oracle:", synth.pose_error(r["pose"], ref["pose"]))
n = min(r["outer_iter"], ref["outer_iter"])
d = [synth.pose_error(r["pass_pose"][k], ref["pass_pose"][k]) for k in range(n)]
first_bad = next((k for k, (a, b) in enumerate(d) if a > 1e-5 or b > 1e-4), None)
print("first pass whose pose differs by more than the tolerance:", first_bad, "lm iters gpu", list(r["pass_lm_iters"][:8]), "oracle", list(ref["pass_lm_iters"][:8]))
print("per-pass diffs (first 10):", [(float("%.1e" % a), float("%.1e" % b)) for a, b in d[:10]])
