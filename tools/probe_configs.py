"""Single-registration device time of every algorithm at the BASELINE config shapes (C1 room 10k, C2 KITTI 120k, C3 NYU 307k)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
cfgs = [("C1 room 10k N=11", synth.room_pair(seed=100, n_points=10_000)), ("C2 KITTI 120k N=20", synth.kitti_pair(0)), ("C3 NYU 307k N=40", synth.nyu_pair(0))]
for name, p in cfgs:
    for algo, an in ((sicp.ALGO_GICP, "GICP"), (sicp.ALGO_SEMANTIC, "SemanticICP"), (sicp.ALGO_EM, "EM-ICP")):
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            if algo == sicp.ALGO_GICP:
                s, t, o = sicp.Cloud(p["src_xyz"]), sicp.Cloud(p["tgt_xyz"]), sicp.default_options(algo, profile=True)
            elif algo == sicp.ALGO_EM:
                s, t, o = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"]), sicp.default_options(algo, cm=p["cm"], profile=True)
            else:
                s = sicp.Cloud(p["src_xyz"], p["src_labels"], layout=sicp.CLOUD_PER_CLASS)
                t = sicp.Cloud(p["tgt_xyz"], p["tgt_labels"], layout=sicp.CLOUD_PER_CLASS)
                o = sicp.default_options(algo, profile=True)
            r = sicp.register(algo, s, t, o, p["init"])
            dt = (time.perf_counter() - t0) * 1e3
            if best is None or dt < best[0]:
                best = (dt, r)
            s.close(); t.close()
        dt, r = best
        rot, trans = synth.pose_error(r["pose"], p["T_gt"])
        st = r["stage_ms"]
        print("%-20s %-12s total %.2f ms (create+register, host clock) | cov %.2f knn %.2f estep %.2f lm %.2f | passes %d lm iters %d | vs ground truth %.1e rad %.1e m" %
              (name, an, dt, st["cov"], st["knn"], st["estep"], st["lm"], r["outer_iter"], r["lm_iters_total"], rot, trans), flush=True)
