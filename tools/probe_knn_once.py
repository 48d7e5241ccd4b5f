"""One cross-kNN (k = 4) and one covariance precompute on pair 0 — the target of `ncu` captures (SICP_LIB selects the build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
p = synth.cached("kitti_pair", 0)
n = len(p["src_xyz"])
dev = torch.device("cuda", 0)
s, t = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
o_idx = torch.empty(n * 4, dtype=torch.int32, device=dev); o_d2 = torch.empty(n * 4, dtype=torch.float32, device=dev)
for _ in range(2):
    sicp.knn_cloud(t, s, 4, o_idx.data_ptr(), o_d2.data_ptr(), pose7=p["T_gt"])
s.precompute(20, 1e-3, p["cm"])
torch.cuda.synchronize()
