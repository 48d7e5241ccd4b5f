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
if os.environ.get("PROBE_WALL"):  # wall clock of a lone registration including cloud creation (slots and graph exec are warm)
    import time, torch
    for i in (1, 2, 3):
        q = synth.kitti_pair(i, n_points=n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a, b = sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])
        r = sicp.register(sicp.ALGO_EM, a, b, opts, q["init"])
        print("lone wall ms", round(1e3 * (time.perf_counter() - t0), 3), "outer", r["outer_iter"], "lm", r["lm_iters_total"])
