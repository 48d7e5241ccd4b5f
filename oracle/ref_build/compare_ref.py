"""Runs oracle/_ref/sicp_ref_runner (the unmodified reference) on every fixture and compares it with the CPU oracle:
final pose within 1e-5 rad / 1e-4 m (the north-star tolerance), the same number of outer passes and the same Ceres iteration
count in every pass.  Exit code 0 = the oracle is pinned to the reference on these inputs; 77 = runner not built."""
import glob
import json
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import semantic_icp_b200 as pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

RUNNER = os.path.join(ROOT, "oracle", "_ref", "sicp_ref_runner")


def read_fixture(path):
    with open(path, "rb") as f:
        magic, ns, nt, N = struct.unpack("<4i", f.read(16))
        assert magic == 0x53494350
        init = np.frombuffer(f.read(56), "<f8")
        cm = np.frombuffer(f.read(8 * N * N), "<f8").reshape(N, N)
        sxyz = np.frombuffer(f.read(12 * ns), "<f4").reshape(ns, 3)
        slab = np.frombuffer(f.read(4 * ns), "<u4")
        txyz = np.frombuffer(f.read(12 * nt), "<f4").reshape(nt, 3)
        tlab = np.frombuffer(f.read(4 * nt), "<u4")
    return dict(init=init, cm=cm, src_xyz=sxyz, src_labels=slab, tgt_xyz=txyz, tgt_labels=tlab)


def main():
    if not os.path.exists(RUNNER):
        print("sicp_ref_runner is not built (PCL / Eigen / Sophus / Ceres missing): parity stays unpinned")
        return 77
    bad = 0
    for fx in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "fixtures", "*.bin"))):
        p = read_fixture(fx)
        for algo in ("gicp", "em", "semantic"):
            ref = json.loads(subprocess.run([RUNNER, fx, algo], capture_output=True, text=True, check=True).stdout)
            if algo == "gicp":
                o = O.align_gicp(p["src_xyz"], p["tgt_xyz"], p["init"])
            elif algo == "em":
                o = O.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["init"])
            else:
                o = O.align_semantic(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["init"])
            rot, tr = pkg.synth.pose_error(np.array(ref["pose7"]), o["pose"])
            same_iters = list(ref["pass_lm_iters"]) == [int(x) for x in o["pass_lm_iters"]]
            same_outer = ref["outer_iter"] in (-1, int(o["outer_iter"]))
            ok = rot < 1e-5 and tr < 1e-4 and same_outer and same_iters
            bad += not ok
            print(f"{os.path.basename(fx)} {algo}: rot {rot:.2e} rad, trans {tr:.2e} m, outer ref {ref['outer_iter']} / oracle {o['outer_iter']}, "
                  f"LM iterations per pass equal: {same_iters} -> {'OK' if ok else 'MISMATCH'}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
