"""Timeline of one batch (SICP_TRACE): per-stage busy intervals of every registration, concurrency and gaps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
conc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pairs = [synth.cached("kitti_pair", i) for i in range(B)]
p = pairs[0]
inits = np.stack([q["init"] for q in pairs])
path = "/tmp/sicp_trace.txt"
for prof in (False, True):
    opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"], profile=prof)
    opts.max_concurrent = conc
    for rep in range(3):
        if os.path.exists(path): os.remove(path)
        if prof: os.environ["SICP_TRACE"] = path
        cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
        t0 = time.perf_counter()
        res = sicp.register_batch(sicp.ALGO_EM, [c[0] for c in cl], [c[1] for c in cl], opts, inits)
        dt = time.perf_counter() - t0
    print("profile", prof, "register_batch %.2f ms" % (dt * 1e3))
rows = np.loadtxt(path)
names = {1: "cov", 2: "knn", 3: "estep", 4: "lm"}
end = rows[:, 3].max()
print("trace span %.2f ms, %d intervals" % (end, len(rows)))
for st, nm in names.items():
    r = rows[rows[:, 1] == st]
    d = r[:, 3] - r[:, 2]
    print("%-6s n=%3d  sum %.2f ms  mean %.3f  max %.3f" % (nm, len(r), d.sum(), d.mean(), d.max()))
# concurrency histogram over time (how many registrations have a kernel in flight)
ts = np.linspace(0, end, 4000)
active = np.zeros_like(ts)
lm_active = np.zeros_like(ts)
for j, st, a, b in rows:
    m = (ts >= a) & (ts < b)
    active += m
    if st == 4: lm_active += m
print("time-weighted: mean kernels in flight %.2f; fraction of time with 0 in flight %.3f; mean LM kernels in flight %.2f" % (active.mean(), (active == 0).mean(), lm_active.mean()))
for j in range(min(B, 4)):
    r = rows[rows[:, 0] == j]
    print("job", j, " ".join("%s[%.2f-%.2f]" % (names[int(s)], a, b) for _, s, a, b in r[np.argsort(r[:, 2])][:14]))
