"""Wall time of single registrations and of a 16-pair batch as a function of SICP_CHUNK (passes per control-block readback)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
pairs = [synth.kitti_pair(i) for i in range(16)]
p = pairs[0]
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
inits = np.stack([q["init"] for q in pairs])
for chunk in sys.argv[1].split(","):
    os.environ["SICP_CHUNK"] = chunk
    s, t = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
    s.precompute(20, 1e-3, p["cm"]); t.precompute(20, 1e-3, p["cm"])
    best = 1e9
    for rep in range(8):
        t0 = time.perf_counter(); r = sicp.register(sicp.ALGO_EM, s, t, opts, p["init"]); best = min(best, time.perf_counter() - t0)
    bb = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
        res = sicp.register_batch(sicp.ALGO_EM, [c[0] for c in cl], [c[1] for c in cl], opts, inits)
        bb = min(bb, time.perf_counter() - t0)
    print("chunk %s: single registration (clouds + covariances cached) %.3f ms, %d passes; batch of 16: %.2f ms -> %.1f reg/s" % (chunk, best * 1e3, r["outer_iter"], bb * 1e3, 16 / bb), flush=True)
