"""Exploratory timing probe (not the bench): stage times of one KITTI-shaped registration + batch throughput."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pairs = [synth.kitti_pair(i, n_points=n) for i in range(B)]
p = pairs[0]
for algo, name in [(sicp.ALGO_EM, "em"), (sicp.ALGO_GICP, "gicp")]:
    for rep in range(3):
        t0 = time.perf_counter()
        if algo == sicp.ALGO_EM:
            src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
            opts = sicp.default_options(algo, cm=p["cm"], profile=True)
        else:
            src, tgt = sicp.Cloud(p["src_xyz"]), sicp.Cloud(p["tgt_xyz"])
            opts = sicp.default_options(algo, profile=True)
        t1 = time.perf_counter()
        r = sicp.register(algo, src, tgt, opts, p["init"])
        t2 = time.perf_counter()
        print(name, rep, "create %.2f ms register %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), "outer", r["outer_iter"], "lm", r["lm_iters_total"],
              "evals", r["lm_evals_total"], "ncorr", r["n_corr_last"], {k: round(v, 3) for k, v in r["stage_ms"].items() if v}, r["stage_launches"],
              "err", synth.pose_error(r["pose"], p["T_gt"]))
# batch throughput, EM
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
for conc in (1, 2, 3, 4, 8):
    opts.max_concurrent = conc
    for rep in range(3):
        t0 = time.perf_counter()
        cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
        t1 = time.perf_counter()
        res = sicp.register_batch(sicp.ALGO_EM, [c[0] for c in cl], [c[1] for c in cl], opts, np.stack([q["init"] for q in pairs]))
        t2 = time.perf_counter()
    print("batch", B, "conc", conc, "create %.2f ms register %.2f ms -> %.1f reg/s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, B / (t2 - t0)), [r["outer_iter"] for r in res])
