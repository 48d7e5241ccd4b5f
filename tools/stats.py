"""Traversal statistics with the instrumented build (make -C semantic-icp_b200 stats)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
sicp.LIB_PATH = os.environ.get("SICP_STATS_LIB") or os.path.join(pkg.PKG_ROOT, "lib", "libsicp_b200_stats.so")
L = sicp.lib()
def stats(reset=True):
    out = (C.c_ulonglong * 8)()
    L.sicp_debug_stats(out, C.c_int(int(reset)))
    return list(out)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
p = synth.kitti_pair(0, n_points=n)
src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
stats()
nw = (n + 31) // 32
src.precompute(20, 1e-3, p["cm"])
s = stats()
print("self kNN k=20: per warp: node expansions %.1f leaf scans %.1f phase2 iters %.1f ; per query insertions %.1f; lanes needing a scanned leaf %.1f, scans with <=8 needers %.0f%%, <=16 %.0f%%" % (s[0] / nw, s[1] / nw, s[3] / nw, s[2] / n, s[4] / max(1, s[1]), 100 * s[5] / max(1, s[1]), 100 * s[6] / max(1, s[1])))
tgt.precompute(20, 1e-3, p["cm"]); stats()
for k in (1, 4):
    for pose, name in ((p["init"], "identity"), (p["T_gt"], "T_gt")):
        sicp.knn(tgt, p["src_xyz"], k, pose7=pose)
        s = stats()
        print("cross kNN k=%d at %s: per warp: node expansions %.1f leaf scans %.1f phase2 iters %.1f ; per query insertions %.1f; lanes needing a scanned leaf %.1f, scans with <=8 needers %.0f%%, <=16 %.0f%%" % (k, name, s[0] / nw, s[1] / nw, s[3] / nw, s[2] / n, s[4] / max(1, s[1]), 100 * s[5] / max(1, s[1]), 100 * s[6] / max(1, s[1])))
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
r = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
cyc = r["lm_cycles"]; ev = r["lm_evals_total"]
print("LM: evals %d, cycles/eval: sweep %.0f wait %.0f control %.0f (reduce %.0f, state load %.0f, lm step %.0f)" % (ev, cyc[0] / ev, cyc[1] / ev, cyc[2] / ev, cyc[3] / ev, cyc[4] / ev, cyc[5] / ev))

# per-warp cost distribution of one search kernel (load balance)
def trace(nwarps):
    buf = (C.c_uint * (2 * nwarps))()
    L.sicp_debug_warp_trace(buf, C.c_int(nwarps))
    a = np.array(buf[:], dtype=np.int64)
    return a[:nwarps], a[nwarps:]
def describe(name, sc, cy):
    q = lambda a, p: np.percentile(a, p)
    print("%s: scans/warp mean %.1f p50 %.0f p90 %.0f p99 %.0f max %d | cycles/warp mean %.0f p50 %.0f p90 %.0f p99 %.0f max %d (sum/max = %.0f warps' worth)" %
          (name, sc.mean(), q(sc, 50), q(sc, 90), q(sc, 99), sc.max(), cy.mean(), q(cy, 50), q(cy, 90), q(cy, 99), cy.max(), cy.sum() / cy.max()))
sicp.knn(tgt, p["src_xyz"], 4, pose7=p["T_gt"])
sc, cy = trace(nw)
describe("cross k=4", sc, cy)
heavy = np.argsort(-cy)[:10]
print("  heaviest warps:", [(int(i), int(sc[i]), int(cy[i])) for i in heavy])
c2 = sicp.Cloud(p["src_xyz"], p["src_labels"]); c2.precompute(20, 1e-3, p["cm"])
sc, cy = trace(nw)
describe("self k=20", sc, cy)
heavy = np.argsort(-cy)[:10]
print("  heaviest warps:", [(int(i), int(sc[i]), int(cy[i])) for i in heavy])

# LM: per-block sweep / wait cycles (is the controller waiting for slow sweeps or for flag / counter propagation?)
blk = (C.c_ulonglong * 1024)()
L.sicp_debug_lm_blocks(blk, C.c_int(1))
r = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
L.sicp_debug_lm_blocks(blk, C.c_int(1))
a = np.array(blk[:], dtype=np.float64).reshape(512, 2) / r["lm_evals_total"]
a = a[: max(1, int(np.count_nonzero(a[:, 0])))]
print("LM per block, cycles/eval: sweep+reduce mean %.0f min %.0f p50 %.0f p90 %.0f max %.0f (block 0: %.0f) | wait-for-pose mean %.0f (block 0 waits for arrivals: %.0f)" %
      (a[1:, 0].mean(), a[1:, 0].min(), np.percentile(a[1:, 0], 50), np.percentile(a[1:, 0], 90), a[1:, 0].max(), a[0, 0], a[1:, 1].mean(), a[0, 1]))
order = np.argsort(-a[:, 0])[:8]
print("  slowest blocks:", [(int(b), int(a[b, 0])) for b in order])
