"""Writes the binary fixtures sicp_ref_runner reads (oracle/_ref/fixtures/*.bin): the SAME seeded inputs as
tests/golden/align_small.json, so that the reference, the oracle and the CUDA path can be compared on identical data."""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import semantic_icp_b200 as pkg  # noqa: E402


def write_fixture(path, p):
    N = int(p["N"]) if "N" in p else p["cm"].shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", 0x53494350, len(p["src_xyz"]), len(p["tgt_xyz"]), N))
        f.write(np.asarray(p["init"], "<f8").tobytes())
        f.write(np.ascontiguousarray(p["cm"], "<f8").tobytes())
        f.write(np.ascontiguousarray(p["src_xyz"], "<f4").tobytes())
        f.write(np.ascontiguousarray(p["src_labels"], "<u4").tobytes())
        f.write(np.ascontiguousarray(p["tgt_xyz"], "<f4").tobytes())
        f.write(np.ascontiguousarray(p["tgt_labels"], "<u4").tobytes())


def main():
    out = os.path.join(ROOT, "oracle", "_ref", "fixtures")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "align_small.json")) as f:
        gold = json.load(f)
    seen = set()
    for c in (gold["cases"] if isinstance(gold, dict) else gold):
        key = (c["seed"], c["n_points"])
        if key in seen:
            continue
        seen.add(key)
        p = pkg.synth.room_pair(seed=c["seed"], n_points=c["n_points"])
        write_fixture(os.path.join(out, f"room_s{c['seed']}_n{c['n_points']}.bin"), p)
    p = pkg.synth.room_pair(seed=100, n_points=10_000)  # configs[0] (exec/test_icp shape)
    write_fixture(os.path.join(out, "room_s100_n10000.bin"), p)
    print("fixtures in", out)


if __name__ == "__main__":
    main()
