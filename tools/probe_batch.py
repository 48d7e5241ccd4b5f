"""Batch throughput probe: registrations/s for combinations of LM grid size (SICP_LM_GRID) and concurrency."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pairs = [synth.cached("kitti_pair", i) for i in range(B)]
p = pairs[0]
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
inits = np.stack([q["init"] for q in pairs])
for grid in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["148", "74", "37"]):
    os.environ["SICP_LM_GRID"] = grid
    for conc in [int(x) for x in (sys.argv[3].split(',') if len(sys.argv) > 3 else ['2','4','6','8'])]:
        opts.max_concurrent = conc
        best = 1e9
        for rep in range(4):
            t0 = time.perf_counter()
            cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
            res = sicp.register_batch(sicp.ALGO_EM, [c[0] for c in cl], [c[1] for c in cl], opts, inits)
            dt = time.perf_counter() - t0
            best = min(best, dt)
        print("lm_grid %s conc %d: %.2f ms per batch of %d -> %.1f reg/s" % (grid, conc, best * 1e3, B, B / best), flush=True)
