"""The reference-compatible C++ facade (semantic-icp_b200/facade/*.h) compiled as C++11 against the minimal
Eigen/Sophus/PCL stand-ins and driven like exec/test_icp.cc / exec/kitti_eval.cc drive the reference.

CPU (-m "not gpu"): the facade compiles and links against libsicp_b200.so, and without a CUDA device it FAILS LOUDLY
(no fallback).  GPU: poses from the three classes agree with the oracle within the north-star tolerance."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROT_TOL, TRANS_TOL = 1e-5, 1e-4


@pytest.fixture(scope="module")
def facade_bin(tmp_path_factory, sicp):
    out = str(tmp_path_factory.mktemp("facade") / "facade_check")
    libdir = os.path.join(ROOT, "semantic-icp_b200", "lib")
    cmd = ["/usr/bin/g++", "-std=c++11", "-O2", "-Wall", "-Werror", f"-I{ROOT}/semantic-icp_b200/facade", f"-I{ROOT}/include",
           os.path.join(ROOT, "tests", "cpp", "facade_check.cc"), "-o", out, f"-L{libdir}", "-lsicp_b200", f"-Wl,-rpath,{libdir}"]
    subprocess.check_call(cmd)
    return out


def _write_input(path, p):
    with open(path, "wb") as f:
        f.write(np.array([len(p["src_xyz"]), len(p["tgt_xyz"]), p["N"]], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(p["src_xyz"], dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(p["src_labels"], dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(p["tgt_xyz"], dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(p["tgt_labels"], dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(p["cm"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(p["init"], dtype=np.float64).tobytes())


def test_facade_compiles_and_fails_loudly_without_gpu(facade_bin, pkg, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; the loud-failure path needs a CPU-only box")
    p = pkg.synth.room_pair(seed=100, n_points=500)
    inp = str(tmp_path / "in.bin")
    _write_input(inp, p)
    r = subprocess.run([facade_bin, inp], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_facade_matches_oracle(facade_bin, pkg, oracle, tmp_path):
    p = pkg.synth.room_pair(seed=100, n_points=10_000)
    inp = str(tmp_path / "in.bin")
    _write_input(inp, p)
    r = subprocess.run([facade_bin, inp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    refs = {
        "gicp": oracle.align_gicp(p["src_xyz"], p["tgt_xyz"], p["init"]),
        "em": oracle.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["init"]),
        "semantic": oracle.align_semantic(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["init"]),
    }
    for name, ref in refs.items():
        rot, trans = pkg.synth.pose_error(np.array(got[name]["pose"]), ref["pose"])
        assert rot < ROT_TOL and trans < TRANS_TOL, (name, rot, trans)
        assert got[name]["outer_iter"] == ref["outer_iter"], name
    # pcl_2_semantic: first-appearance class order; per-class covariances and kd-tree through the public maps
    cl, cs, order = oracle.label_split(p["src_labels"])
    assert got["labels_first_appearance"] == [int(x) for x in cl]
    cls0 = p["src_xyz"][order[cs[0]:cs[1]]]
    cov = oracle.covariances(cls0, 20, 1e-3)["cov"][0]
    assert np.max(np.abs(np.array(got["class0_cov0"]).reshape(3, 3) - cov)) <= 1e-12
    idx, d2 = oracle.knn(cls0, cls0[:1], 3)
    assert got["class0_knn3"]["found"] == 3 and got["class0_knn3"]["idx"] == [int(i) for i in idx[0]]
    assert np.array_equal(np.array(got["class0_knn3"]["d2"], dtype=np.float32), d2[0])
    # finalCloud = float-matrix transform of the source (impl/gicp.hpp:166-171)
    M = pkg.synth.pose7_matrix(np.array(got["gicp"]["pose"])).astype(np.float32)
    x = p["src_xyz"][0]
    exp = np.array([M[r, 0] * x[0] + M[r, 1] * x[1] + M[r, 2] * x[2] + M[r, 3] for r in range(3)], dtype=np.float32)
    assert np.allclose(np.array(got["gicp_final0"], dtype=np.float32), exp, atol=1e-5)
    assert got["gicp_src_cov_n"] == len(p["src_xyz"])
    assert got["fused_n"] == len(p["src_xyz"]) and got["fused_same_as_input"] > 0.8 * len(p["src_xyz"])


def test_compat_shims_against_golden_se3(tmp_path):
    """The Eigen / Sophus / PCL stand-ins (facade/compat) on the CPU: layouts, transformPointCloud, SE(3) golden vectors."""
    with open(os.path.join(ROOT, "tests", "golden", "se3.json")) as f:
        items = json.load(f)
    flat = tmp_path / "se3.txt"
    with open(flat, "w") as f:
        for it in items:
            row = list(it["delta"]) + list(it["pose7"]) + list(it["delta_b"]) + list(it["pose7_ab"]) + list(it["pose7_inv"])
            f.write(" ".join(repr(float(x)) for x in row) + "\n")
    exe = str(tmp_path / "shim_check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-DSICP_FACADE_FORCE_SHIMS", f"-I{ROOT}/semantic-icp_b200/facade",
                           f"-I{ROOT}/include", os.path.join(ROOT, "tests", "cpp", "shim_check.cc"), "-o", exe])
    r = subprocess.run([exe, str(flat), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "shim_check OK" in r.stdout


def _write_pcd(path, xyz, labels, binary):
    n = len(xyz)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z label\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {'binary' if binary else 'ascii'}\n")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("l", "<u4")])
            rec["x"], rec["y"], rec["z"], rec["l"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], labels
            f.write(rec.tobytes())
        else:
            for p, l in zip(xyz, labels):
                f.write(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r} {int(l)}\n".encode())


@pytest.mark.gpu
def test_example_driver_pcd_to_metrics(pkg, oracle, sicp, tmp_path):
    """examples/pair_eval.cc: PCD in -> range filter -> EM-ICP + GICP through the facade -> SE(3) errors + label agreement."""
    p = pkg.synth.room_pair(seed=31, n_points=6000)
    _write_pcd(tmp_path / "a.pcd", p["src_xyz"], p["src_labels"], binary=True)
    _write_pcd(tmp_path / "b.pcd", p["tgt_xyz"], p["tgt_labels"], binary=False)
    np.savetxt(tmp_path / "cm.txt", p["cm"], fmt="%.17g")
    exe = str(tmp_path / "pair_eval")
    libdir = os.path.join(ROOT, "semantic-icp_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O2", "-Wall", "-Werror", f"-I{ROOT}/semantic-icp_b200/facade", f"-I{ROOT}/include",
                           os.path.join(ROOT, "examples", "pair_eval.cc"), "-o", exe, f"-L{libdir}", "-lsicp_b200", f"-Wl,-rpath,{libdir}"])
    rng_m = 5.5  # drops the far corners of the room
    r = subprocess.run([exe, str(tmp_path / "a.pcd"), str(tmp_path / "b.pcd"), str(tmp_path / "cm.txt"), str(rng_m)] + [repr(float(v)) for v in p["T_gt"]],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    ka, kb = oracle.filter_range(p["src_xyz"], rng_m), oracle.filter_range(p["tgt_xyz"], rng_m)
    assert got["n_source"] == len(ka) < len(p["src_xyz"]) and got["n_target"] == len(kb)
    sx, sl, tx, tl = p["src_xyz"][ka], p["src_labels"][ka], p["tgt_xyz"][kb], p["tgt_labels"][kb]
    ident = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    for name, ref in (("em", oracle.align_em(sx, sl, tx, tl, p["cm"], ident)), ("gicp", oracle.align_gicp(sx, tx, ident))):
        rot, trans = pkg.synth.pose_error(np.array(got[name]["pose"]), ref["pose"])
        assert rot < ROT_TOL and trans < TRANS_TOL and got[name]["outer_iter"] == ref["outer_iter"], (name, rot, trans)
        assert np.allclose(got[name + "_error"], oracle.pose_errors(p["T_gt"], np.array(got[name]["pose"])), rtol=1e-5, atol=1e-12)
    la = got["label_agreement"]
    assert 0.5 < la["accuracy"] <= 1.0 and la["pairs"] == len(ka) and la["mean_distance"] < 0.3
