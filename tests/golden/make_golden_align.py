#!/usr/bin/env python
"""Generates tests/golden/align_small.json with tests/golden/numpy_reference.py (an independent numpy transcription of
the reference's align() loops; never touches the oracle).  Inputs are the seeded synthetic room pairs of
semantic-icp_b200/python/synth.py at sizes pure numpy finishes in minutes.
Run: python tests/golden/make_golden_align.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import numpy_reference as ref  # noqa: E402
import semantic_icp_b200 as pkg  # noqa: E402

cases = []
for algo, seed, n in (("gicp", 41, 400), ("em", 42, 350), ("semantic", 43, 1700), ("gicp", 44, 500)):
    p = pkg.synth.room_pair(seed=seed, n_points=n, N=3 if algo == "semantic" else 11)
    if algo == "gicp":
        pose, outer, lm = ref.align_gicp(p["src_xyz"], p["tgt_xyz"], p["init"])
    elif algo == "em":
        pose, outer, lm = ref.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["init"])
    else:
        pose, outer, lm = ref.align_semantic(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["init"])
    print(algo, seed, n, "outer", outer, "lm", lm, "err vs gt", pkg.synth.pose_error(pose, p["T_gt"]), flush=True)
    cases.append(dict(algo=algo, seed=seed, n_points=n, N=3 if algo == "semantic" else 11, pose7=[float(v) for v in pose], outer_iter=int(outer), lm_iters=[int(v) for v in lm]))
with open(os.path.join(HERE, "align_small.json"), "w") as f:
    json.dump(cases, f, indent=1)
print("wrote align_small.json")

# getFusedLabels on a small pair (labels + the margin between the two best classes, to recognise exact ties)
p = pkg.synth.room_pair(seed=45, n_points=400)
lab, margin = ref.fused_labels(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["T_gt"])
with open(os.path.join(HERE, "fused_small.json"), "w") as f:
    json.dump(dict(seed=45, n_points=400, labels=[int(x) for x in lab], margin=[float(x) for x in margin]), f)
print("wrote fused_small.json")
