"""Small end-to-end run of every entry point, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
p = synth.room_pair(seed=5, n_points=3000)
src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
if os.environ.get("SAN_STAGES") == "search":  # searches, covariances, first-pass correspondences only (minutes faster under memcheck)
    for k in (1, 4, 20):
        print("knn", k, sicp.knn(tgt, p["src_xyz"], k, pose7=p["init"])[0].shape)
    src.precompute(20, 1e-3, p["cm"])
    c2 = sicp.Cloud(p["src_xyz"], p["src_labels"], layout=sicp.CLOUD_PER_CLASS)
    c2.precompute(20, 1e-3)
    print("getters", src.normals().shape, src.label_vectors().shape, src.self_neighbours().shape, c2.covariances().shape)
    idx, w, d2 = sicp.correspondences(sicp.ALGO_EM, src, tgt, opts, p["init"])
    print("corr", int((idx >= 0).sum()), float(w.sum()))
    sys.exit(0)
r = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
print("em", r["outer_iter"], r["lm_iters_total"])
g = sicp.register(sicp.ALGO_GICP, sicp.Cloud(p["src_xyz"]), sicp.Cloud(p["tgt_xyz"]), sicp.default_options(sicp.ALGO_GICP), p["init"])
print("gicp", g["outer_iter"])
ps = synth.room_pair(seed=6, n_points=6000)
s2 = sicp.Cloud(ps["src_xyz"], ps["src_labels"], layout=sicp.CLOUD_PER_CLASS)
t2 = sicp.Cloud(ps["tgt_xyz"], ps["tgt_labels"], layout=sicp.CLOUD_PER_CLASS)
s = sicp.register(sicp.ALGO_SEMANTIC, s2, t2, sicp.default_options(sicp.ALGO_SEMANTIC), ps["init"])
print("semantic", s["outer_iter"])
b = sicp.register_batch(sicp.ALGO_EM, [src] * 3, [tgt] * 3, opts, np.stack([p["init"]] * 3))
print("batch", [x["outer_iter"] for x in b])
print("knn", sicp.knn(tgt, p["src_xyz"][:100], 20)[0].shape, "fused", sicp.fused_labels(src, tgt, opts, r["pose"]).shape)
print("metrics", sicp.label_agreement(src, tgt, p["N"] + 1, pose7=r["pose"])["total"], len(sicp.filter_range(p["src_xyz"], 4.0)), src.transform_f32(r["pose"]).shape)
src.precompute(20, 1e-3, p["cm"])
print("getters", src.normals().shape, src.covariances().shape, src.label_vectors().shape, src.self_neighbours().shape)
# round 2: pass-by-pass path beside the graph loop, pose averaging / fusion, tiny epsilon (the denormal-range Probability gate)
os.environ["SICP_GRAPH"] = "0"
r0 = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
os.environ["SICP_GRAPH"] = "1"
print("pass-by-pass", r0["outer_iter"], np.array_equal(r0["pose"], r["pose"]))
from oracle import oracle as O
poses = np.array([O.se3_plus(r["pose"], np.random.default_rng(i).normal(scale=0.02, size=6)) for i in range(7)])
print("mean", sicp.iterative_mean(poses, 50)[1], "fusion", sicp.pose_fusion(poses, np.array([np.eye(6) * 1e-4] * 7), r["pose"])[1])
o6 = sicp.default_options(sicp.ALGO_EM, cm=p["cm"], epsilon=1e-6)
idx, w, d2 = sicp.correspondences(sicp.ALGO_EM, src, tgt, o6, np.array([0, 0, 0, 1.0, 0, 0, 0.9]))
print("gate zeroed", int(((idx >= 0) & (w == 0)).sum()))
