#!/usr/bin/env python
"""Generates the golden fixtures in this directory WITHOUT using the oracle (numpy / scipy / OpenCV only).

The reference ships no tests and no golden vectors, and none of its dependencies can be built here,
so the fixtures are DERIVED: each one transcribes a reference formula independently of oracle/sicp_oracle.cpp, or
uses an independent library as a second opinion.

  cost_function.json  gicp_cost_function.h:27-73,98-176 transcribed to numpy on the constants of
                      exec/test_gradient.cc:32-50 (the only numeric fixture in the reference): residual and 1x7
                      Jacobian at identity and at 10 seeded poses; 6-dof Jacobian by central differences of
                      r(T*exp(delta)) with scipy's exp map.
  knn.npz             exact kNN by the FP32 brute-force definition ((d2, index) order) on a seeded cloud, checked
                      here against OpenCV's FLANN-lineage KDTREE_SINGLE exact search.
  se3.json            exp / log / composition of SE(3) elements from scipy.spatial.transform + closed-form V matrix.
Run: python tests/golden/make_golden.py   (writes next to this file)
"""
import json
import os

import numpy as np
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))

PS = np.array([7.96094, -5.25134, 24.2516], dtype=np.float32).astype(np.float64)   # test_gradient.cc:32
PT = np.array([17.73844, -5.16017, 14.3069], dtype=np.float32).astype(np.float64)  # test_gradient.cc:34
CS = np.array([[0.674143, 0.460412, 0.085842], [0.460412, 0.349471, -0.121288], [0.085842, -0.121288, 0.977386]])
CT = CS.copy()
CT[0, 0] = 0.074143                                                                 # test_gradient.cc:48


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def evaluate(pose7):
    """gicp_cost_function.h:27-73 — residual and 1x7 Jacobian [qx,qy,qz,qw,tx,ty,tz]."""
    q, t = pose7[:4], pose7[4:]
    R = quat_R(q)
    M = np.linalg.inv(CT + R @ CS @ R.T)
    res = PT - (R @ PS + t)
    dT = M @ res
    r = float(res @ dT)
    Ta = np.linalg.inv(CT.T + R @ CS.T @ R.T)
    tb, tc = M @ res, Ta @ res
    dR = -(np.outer(tb, PS) + np.outer(tc, res @ Ta @ R @ CS.T) + np.outer(tb, res @ M @ R @ CS) + np.outer(tc, PS))
    x, y, z, w = q
    tx, ty, tz, tw = 2 * x, 2 * y, 2 * z, 2 * w
    mfx, mfy, mfz, mtw = -2 * tx, -2 * ty, -2 * tz, -tw
    dRdw = np.array([[0, -tz, ty], [tz, 0, -tx], [-ty, tx, 0]])
    dRdx = np.array([[0, ty, tz], [ty, mfx, mtw], [tz, tw, mfx]])
    dRdy = np.array([[mfy, tx, tw], [tx, 0, tz], [mtw, tz, mfy]])
    dRdz = np.array([[mfz, mtw, tx], [tw, mfz, ty], [tx, ty, 0]])
    jac = np.array([np.trace(dR.T @ dRdx), np.trace(dR.T @ dRdy), np.trace(dR.T @ dRdz), np.trace(dR.T @ dRdw), *(-2.0 * dT)])
    return r, jac


def se3_exp(d):
    ups, om = d[:3], d[3:]
    th = np.linalg.norm(om)
    Rm = Rotation.from_rotvec(om).as_matrix()
    Om = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if th < 1e-9:
        V = np.eye(3) + 0.5 * Om
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * Om + (th - np.sin(th)) / th**3 * Om @ Om
    return Rm, V @ ups


def pose7_of(Rm, t):
    q = Rotation.from_matrix(Rm).as_quat()  # x,y,z,w
    if q[3] < 0:
        q = -q
    return np.concatenate([q, t])


def residual_at(Rm, t):
    M = np.linalg.inv(CT + Rm @ CS @ Rm.T)
    res = PT - (Rm @ PS + t)
    return float(res @ M @ res)


def main():
    rng = np.random.default_rng(20240517)
    cases = []
    poses = [np.array([0, 0, 0, 1, 0, 0, 0.0])]
    for _ in range(10):
        Rm, t = se3_exp(np.concatenate([rng.uniform(-2, 2, 3), rng.normal(size=3)]))
        poses.append(pose7_of(Rm, t))
    for p in poses:
        r, j7 = evaluate(p)
        Rm, t = quat_R(p[:4]), p[4:]
        j6 = np.zeros(6)
        h = 1e-6
        for a in range(6):
            e = np.zeros(6)
            e[a] = h
            Rp, tp = se3_exp(e)
            Rn, tn = se3_exp(-e)
            j6[a] = (residual_at(Rm @ Rp, t + Rm @ tp) - residual_at(Rm @ Rn, t + Rm @ tn)) / (2 * h)
        cases.append(dict(pose7=p.tolist(), residual=r, jac7=j7.tolist(), jac6_numeric=j6.tolist()))
    with open(os.path.join(HERE, "cost_function.json"), "w") as f:
        json.dump(dict(ps=PS.tolist(), pt=PT.tolist(), cs=CS.tolist(), ct=CT.tolist(), cases=cases,
                       survey_known_answer=dict(residual=-200.583920753427,
                                                jac7=[2113.425298320, 1595.374239304, -298.8699546585, 0, 41.29076878877,
                                                      -54.74467097423, -0.2451985054000])), f, indent=1)

    # ---- kNN by definition
    tgt = (rng.normal(size=(3000, 3)) * np.array([20, 20, 2])).astype(np.float32)
    qry = (rng.normal(size=(300, 3)) * np.array([20, 20, 2])).astype(np.float32)
    k = 20
    idx = np.zeros((len(qry), k), dtype=np.int32)
    d2o = np.zeros((len(qry), k), dtype=np.float32)
    for i, q in enumerate(qry):
        d = (q - tgt).astype(np.float32)
        d2 = (d[:, 0] * d[:, 0]).astype(np.float32)
        d2 = (d2 + (d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
        d2 = (d2 + (d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
        order = np.lexsort((np.arange(len(tgt)), d2))[:k]
        idx[i], d2o[i] = order, d2[order]
    try:
        import cv2

        fl = cv2.flann_Index(tgt, dict(algorithm=4, leaf_max_size=15))
        fi, fd = fl.knnSearch(qry, k, params=dict(checks=-1, eps=0.0, sorted=True))
        assert np.array_equal(fi, idx) and np.array_equal(fd, d2o), "OpenCV FLANN disagrees with the brute-force definition"
        flann_checked = True
    except ImportError:
        flann_checked = False
    np.savez_compressed(os.path.join(HERE, "knn.npz"), tgt=tgt, qry=qry, idx=idx, d2=d2o, flann_checked=flann_checked)

    # ---- SE(3)
    items = []
    for _ in range(12):
        d = np.concatenate([rng.normal(size=3), rng.uniform(-1.5, 1.5, 3)])
        Rm, t = se3_exp(d)
        d2 = np.concatenate([rng.normal(size=3), rng.uniform(-1.5, 1.5, 3)])
        R2, t2 = se3_exp(d2)
        items.append(dict(delta=d.tolist(), pose7=pose7_of(Rm, t).tolist(), delta_b=d2.tolist(),
                          pose7_ab=pose7_of(Rm @ R2, t + Rm @ t2).tolist(), pose7_inv=pose7_of(Rm.T, -Rm.T @ t).tolist()))
    with open(os.path.join(HERE, "se3.json"), "w") as f:
        json.dump(items, f, indent=1)
    print("golden fixtures written; flann cross-check:", flann_checked)


if __name__ == "__main__":
    main()
