"""Batch / lone throughput sweep over the library's tuning knobs (environment variables read per call).

  python tools/sweep.py B "VARIANT:GRID:CONC[:GRAPH[:REUSE]],..." [reps]      e.g.  16 "0:37:8,1:74:8:0,2:74:12:1:1"
  VARIANT = LM kernel shape (SICP_LM_VARIANT: 0 full registers, 1 128-thread CTAs, 2/3/4 = 176/160/144 registers), GRID = CTAs of a batch solve
  (SICP_LM_GRID), CONC = registrations in flight,
  GRAPH = device-resident outer loop on/off (SICP_GRAPH), REUSE = graph exec updated in place (SICP_GRAPH_REUSE).
Prints registrations/s (best and median of `reps` batches; clouds are created from host arrays inside the timed span, like
the e2e leg of bench.py) and, first, the wall time of one lone registration.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import semantic_icp_b200 as pkg
sicp, synth = pkg.sicp, pkg.synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfgs = (sys.argv[2] if len(sys.argv) > 2 else "0:37:8").split(",")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
npts = int(os.environ.get("NPTS", "120000"))
pairs = [synth.cached("kitti_pair", i, n_points=npts) for i in range(B)]
p = pairs[0]
opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
inits = np.stack([q["init"] for q in pairs])
tag = os.environ.get("SICP_LIB", "default").split("/")[-1] + " graph=" + os.environ.get("SICP_GRAPH", "1") + " reuse=" + os.environ.get("SICP_GRAPH_REUSE", "-")
# lone registration wall time (clouds + register)
ts = []
for rep in range(6):
    t0 = time.perf_counter()
    s, t = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
    r = sicp.register(sicp.ALGO_EM, s, t, opts, p["init"])
    ts.append(time.perf_counter() - t0)
    s.close(); t.close()
print(f"[{tag}] lone registration wall: best {min(ts)*1e3:.2f} ms median {sorted(ts)[len(ts)//2]*1e3:.2f} ms  outer {r['outer_iter']} lm {r['lm_iters_total']}", flush=True)
# marginal cost of the stages in a batch: the same batch with (a) clouds built outside the timed span, (b) clouds built and
# covariances / label vectors precomputed outside it
if os.environ.get("STAGES", "1") != "0":
    opts.max_concurrent = 8
    for mode in ("all inside", "clouds outside", "clouds + precompute outside"):
        ts = []
        for rep in range(reps + 1):
            if mode == "all inside":
                t0 = time.perf_counter()
            cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
            if mode == "clouds + precompute outside":
                for x in cl:
                    x[0].precompute(20, 1e-3, p["cm"]); x[1].precompute(20, 1e-3, p["cm"])
            if mode != "all inside":
                sicp.knn(cl[-1][1], p["src_xyz"][:1], 1)  # synchronous: drains the stream the clouds were built on
                t0 = time.perf_counter()
            res = sicp.register_batch(sicp.ALGO_EM, [x[0] for x in cl], [x[1] for x in cl], opts, inits)
            dt = time.perf_counter() - t0
            for x in cl:
                x[0].close(); x[1].close()
            if rep:
                ts.append(dt)
        print(f"[{tag}] batch of {B}, {mode}: best {min(ts)*1e3:.2f} ms ({min(ts)*1e3/B:.3f} ms per registration)", flush=True)
ref = None
for cfg in cfgs:
    f = cfg.split(":") + ["1", "1"]
    v, g, c = f[0], f[1], f[2]
    os.environ["SICP_LM_VARIANT"], os.environ["SICP_LM_GRID"], os.environ["SICP_GRAPH"], os.environ["SICP_GRAPH_REUSE"] = v, g, f[3], f[4]
    opts.max_concurrent = int(c)
    ts, tc = [], []
    for rep in range(reps + 1):
        t0 = time.perf_counter()
        cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
        t1 = time.perf_counter()
        res = sicp.register_batch(sicp.ALGO_EM, [x[0] for x in cl], [x[1] for x in cl], opts, inits)
        dt = time.perf_counter() - t0
        for x in cl:
            x[0].close(); x[1].close()
        if rep:
            ts.append(dt); tc.append(t1 - t0)
    sig = [(r["outer_iter"], r["lm_iters_total"]) for r in res]
    if ref is None:
        ref = sig
    print(f"[{tag}] variant {v} grid {g} conc {c} graph {f[3]} reuse {f[4]}: best {B/min(ts):.1f} median {B/sorted(ts)[len(ts)//2]:.1f} reg/s  ({min(ts)*1e3:.2f} ms / {B}; host-side cloud creation {min(tc)*1e3:.2f} ms of it)"
          + ("" if sig == ref else "  !! pass/iteration counts differ from the first config"), flush=True)
