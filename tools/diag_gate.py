"""Diagnostic for the bool-Probability gate at eps = 1e-6: which weights does the GPU zero that the oracle keeps (or vice versa)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
from oracle import oracle as O
sicp = pkg.sicp
p = pkg.synth.room_pair(seed=100, n_points=10_000)
eps = 1e-6
init = np.array([0, 0, 0, 1.0, 0, 0, 0.9])
src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"], epsilon=eps)
idx, w, d2 = sicp.correspondences(sicp.ALGO_EM, src, tgt, opts, init)
ref = O.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], init, eps=eps)
r_idx, r_w = ref["corr0"].reshape(idx.shape), ref["w0"].reshape(w.shape)
zg, zr = (idx >= 0) & (w == 0), (r_idx >= 0) & (r_w == 0)
print("idx equal", np.array_equal(idx, r_idx), "zeros gpu", zg.sum(), "oracle", zr.sum(), "gpu-only", (zg & ~zr).sum(), "oracle-only", (zr & ~zg).sum())
ns, nt = O.covariances(p["src_xyz"], 20, eps)["normals"], O.covariances(p["tgt_xyz"], 20, eps)["normals"]
kappa = 1 - eps
for (i, c) in list(zip(*np.nonzero(zg != zr)))[:30]:
    u, m = nt[idx[i, c]], ns[i]
    d = p["tgt_xyz"][idx[i, c]].astype(np.float64) - (p["src_xyz"][i].astype(np.float64) + init[4:])
    cc, pp, qq = u @ m, u @ d, m @ d
    a, be = 1 - 0.5 * kappa, 0.5 * kappa * cc
    idet = 1 / ((a - be) * (a + be))
    b = 0.5 * d + 0.25 * kappa * ((a * pp + be * qq) * idet * u + (be * pp + a * qq) * idet * m)
    mahal = -0.5 * d @ b
    det = 2 * ((2 - kappa) - kappa * cc) * ((2 - kappa) + kappa * cc)
    print(i, c, "gpu w", w[i, c], "oracle w", r_w[i, c], "mahal", mahal, "pow", ((2 * np.pi) ** 3 * det) ** -0.5, "exp", np.exp(mahal))
