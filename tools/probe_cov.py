"""cov-stage time of one KITTI-shaped EM registration (A/B of library builds through SICP_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
p = synth.cached("kitti_pair", 0)
best = {}
for rep in range(5):
    s, t = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
    r = sicp.register(sicp.ALGO_EM, s, t, sicp.default_options(sicp.ALGO_EM, cm=p["cm"], profile=True), p["init"])
    for k, v in r["stage_ms"].items():
        if v: best[k] = min(best.get(k, 1e9), v)
    s.close(); t.close()
print(os.environ.get("SICP_LIB", "default"), {k: round(v, 3) for k, v in best.items()})
