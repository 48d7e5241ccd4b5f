"""Exact-kNN kernel rates of one library build (SICP_LIB): queries/s for k = 1, 4, 20 (120k transformed source points against
the 120k-point target, device-resident) and the covariance precompute time of one cloud, CUDA-event timed, best of 5."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
p = synth.cached("kitti_pair", 0)
n = len(p["src_xyz"])
dev = torch.device("cuda", 0)
s, t = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
out = {}
for k in (1, 4, 20):
    o_idx = torch.empty(n * k, dtype=torch.int32, device=dev); o_d2 = torch.empty(n * k, dtype=torch.float32, device=dev)
    best = 1e9
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            sicp.knn_cloud(t, s, k, o_idx.data_ptr(), o_d2.data_ptr(), pose7=p["T_gt"])
        e1.record(); e1.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1) / 5)
    out[f"k{k}_Mq/s"] = round(n / best / 1e3, 1)
best = 1e9
for rep in range(6):
    c = sicp.Cloud(p["src_xyz"], p["src_labels"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c.precompute(20, 1e-3, p["cm"]); e1.record(); e1.synchronize()
    if rep: best = min(best, e0.elapsed_time(e1))
    c.close()
out["precompute_one_cloud_ms"] = round(best, 4)
print(os.environ.get("SICP_LIB", "default").split("/")[-1], out)
