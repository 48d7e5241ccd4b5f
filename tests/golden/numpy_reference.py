"""An INDEPENDENT numpy transcription of the reference's three align() loops, used only to generate
tests/golden/align_small.json (make_golden_align.py).  It shares no code with oracle/sicp_oracle.cpp and takes
different routes wherever the mathematics allows (numpy.linalg.inv / eigh / solve instead of hand-written 3x3 and 6x6
routines, the closed-form 6-dof Jacobian instead of the 1x7 quaternion chain, scipy Rotation for exp/log), so agreement
of the two on final poses and pass counts pins the oracle's control flow: correspondence gating, class rules, loss
composition, Ceres-style trust-region rules, outer stopping rule.

Reference lines restated: impl/gicp.hpp:29-175,177-239; impl/semantic_icp.hpp:27-166; impl/em_icp.hpp:24-200,270-344;
gicp_cost_function.h:27-87; local_parameterization_se3.h:17-24; sqloss.h:11-19; Ceres trust-region rules as in
SURVEY.md Appendix C.3.  Pure numpy: only for small clouds.
"""
import numpy as np
from scipy.spatial.transform import Rotation

EPS_SOPHUS = 1e-10


# ---------------------------------------------------------------- SE(3): pose7 = [qx,qy,qz,qw,tx,ty,tz]
def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def V_matrix(om):
    th = np.linalg.norm(om)
    O = hat(om)
    if th < 1e-8:
        return np.eye(3) + 0.5 * O + O @ O / 6.0
    return np.eye(3) + (1 - np.cos(th)) / th**2 * O + (th - np.sin(th)) / th**3 * (O @ O)


def se3_exp(d):
    R = Rotation.from_rotvec(d[3:]).as_matrix()
    return R, V_matrix(d[3:]) @ d[:3]


def se3_log(R, t):
    om = Rotation.from_matrix(R).as_rotvec()
    return np.concatenate([np.linalg.solve(V_matrix(om), t), om])


def pose7_to_Rt(p):
    return Rotation.from_quat(p[:4]).as_matrix(), np.asarray(p[4:], dtype=np.float64)


def Rt_to_pose7(R, t):
    q = Rotation.from_matrix(R).as_quat()
    if q[3] < 0:
        q = -q
    return np.concatenate([q, t])


# ---------------------------------------------------------------- exact kNN by the FP32 brute-force definition
def knn(tgt, q, k):
    tgt, q = tgt.astype(np.float32), q.astype(np.float32)
    idx = np.full((len(q), k), -1, dtype=np.int64)
    d2o = np.full((len(q), k), np.inf, dtype=np.float32)
    for i, p in enumerate(q):
        d = tgt - p
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]          # float32 throughout, left to right
        order = np.lexsort((np.arange(len(tgt)), d2))[:k]
        idx[i, : len(order)], d2o[i, : len(order)] = order, d2[order]
    return idx, d2o


def transform_points(R, t, xyz):  # pcl::transformPointCloud<PointT,double>: f64 math, f32 store
    x = xyz.astype(np.float64)
    return np.stack([((R[r, 0] * x[:, 0] + R[r, 1] * x[:, 1]) + R[r, 2] * x[:, 2]) + t[r] for r in range(3)], axis=1).astype(np.float32)


# ---------------------------------------------------------------- covariances (impl/gicp.hpp:177-239, em_icp.hpp:270-344)
def covariances(xyz, k, eps, labels=None, N=0):
    nn, _ = knn(xyz, xyz, k)
    n = len(xyz)
    covs = np.zeros((n, 3, 3))
    dist = np.zeros((n, N))
    for i in range(n):
        p = xyz[nn[i][nn[i] >= 0]].astype(np.float32)
        mean = p.astype(np.float64).sum(0) / k                                        # divisor k even if fewer were found
        c = np.zeros((3, 3))
        for a in range(3):
            for b in range(a + 1):
                c[a, b] = c[b, a] = (p[:, a] * p[:, b]).astype(np.float64).sum() / k - mean[a] * mean[b]   # f32 products
        w, Vv = np.linalg.eigh(c)
        nrm = Vv[:, np.argmin(np.abs(w))]                                              # JacobiSVD orders by |lambda|
        covs[i] = np.eye(3) - (1 - eps) * np.outer(nrm, nrm)                           # U diag(1,1,eps) U^T
        if N:
            for j in nn[i][nn[i] >= 0]:
                dist[i, labels[j] - 1] += 1.0 / k
    return covs, dist


# ---------------------------------------------------------------- residual, Jacobian, losses
def residual_and_jac(R, t, ps, pt, cs, ct):
    M = np.linalg.inv(ct + R @ cs @ R.T)
    d = pt - (R @ ps + t)
    b = M @ d
    r = float(d @ b)
    c = R.T @ b
    return r, np.concatenate([-2 * c, 2 * np.cross(c, ps + cs @ c)])                 # SURVEY §8(c) closed form


def probability_is_nonzero(R, t, ps, pt, cs, ct):  # gicp_cost_function.h:75-87 converted to bool
    C = ct + R @ cs @ R.T
    d = pt - (R @ ps + t)
    dens = np.linalg.det(2 * np.pi * C) ** -0.5 * np.exp(-0.5 * d @ np.linalg.solve(C, d))
    return dens != 0.0


def loss(kind, w, s):
    """(rho, rho') at s: kind 0 Composed(Cauchy(3), SQ), 1 Cauchy(1.5), 2 Composed(Scaled(Cauchy(3), w), SQ)."""
    if kind == 1:
        return 2.25 * np.log1p(s / 2.25), 1.0 / (1.0 + s / 2.25)
    g = np.sqrt(s + np.finfo(np.float64).eps)
    f0, f1 = 9.0 * np.log1p(g / 9.0), 1.0 / (1.0 + g / 9.0)
    if kind == 2:
        f0, f1 = w * f0, w * f1
    return f0, f1 / (2.0 * g)


def evaluate(res, kind, R, t):
    cost, H, g = 0.0, np.zeros((6, 6)), np.zeros(6)
    for (ps, pt, cs, ct, w) in res:
        r, J = residual_and_jac(R, t, ps, pt, cs, ct)
        rho0, rho1 = loss(kind, w, r * r)
        cost += 0.5 * rho0
        sr = np.sqrt(rho1)                                                             # Ceres corrector, rho'' <= 0
        J, rc = sr * J, sr * r
        H += np.outer(J, J)
        g += J * rc
    return cost, g, H


def plus(R, t, delta):  # local_parameterization_se3.h:22 : T * exp(delta)
    Re, te = se3_exp(delta)
    return R @ Re, t + R @ te


# ---------------------------------------------------------------- Ceres-style trust-region LM
def lm_solve(res, kind, R, t):
    if not res:
        return R, t, 0
    tol = 0.1 * EPS_SOPHUS
    cost, g, H = evaluate(res, kind, R, t)
    scale = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    radius, dec, reuse, last_ok, invalid, it = 1e4, 2.0, False, True, 0, 0
    diag = None

    def gmax(R, t, g):
        Rp, tp = plus(R, t, -g)
        return np.max(np.abs(Rt_to_pose7(R, t) - Rt_to_pose7(Rp, tp)))

    gm, xn = gmax(R, t, g), np.linalg.norm(Rt_to_pose7(R, t))
    while True:
        if it >= 400 or (last_ok and gm <= tol) or radius <= 1e-32:
            break
        it += 1
        Hs, gs = H * np.outer(scale, scale), g * scale
        if not reuse:
            diag = np.clip(np.diag(Hs), 1e-6, 1e32)
        reuse = True
        try:
            y = np.linalg.solve(Hs + np.diag(diag / radius), gs)
            ok = np.all(np.linalg.eigvalsh(Hs + np.diag(diag / radius)) > 0)
        except np.linalg.LinAlgError:
            ok = False
        model = (y @ gs - 0.5 * y @ Hs @ y) if ok else -1.0
        if not ok or not model > 0:
            invalid += 1
            if invalid >= 5:
                break
            radius, dec, last_ok = radius / dec, dec * 2, False
            continue
        invalid = 0
        Rc, tc = plus(R, t, -y * scale)
        if np.linalg.norm(Rt_to_pose7(R, t) - Rt_to_pose7(Rc, tc)) <= 1e-8 * (xn + 1e-8):
            break                                                                       # parameter tolerance: candidate not applied
        ccost, cg, cH = evaluate(res, kind, Rc, tc)
        if abs(cost - ccost) <= tol * cost:
            break                                                                       # function tolerance: candidate not applied
        q = (cost - ccost) / model
        if q > 1e-3:
            R, t, cost, g, H = Rc, tc, ccost, cg, cH
            gm, xn = gmax(R, t, g), np.linalg.norm(Rt_to_pose7(R, t))
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2 * q - 1) ** 3))
            dec, reuse, last_ok = 2.0, False, True
        else:
            radius, dec, reuse, last_ok = radius / dec, dec * 2, True, False
    return R, t, it


# ---------------------------------------------------------------- the three align() loops
def _outer(build_residuals, kind, init7, mse_stop, cap, semantic=False):
    R, t = pose7_to_Rt(np.asarray(init7, dtype=np.float64))
    count, lm_iters = 0, []
    while True:
        if semantic:
            count += 1                                                                  # semantic_icp.hpp:47
        res = build_residuals(R, t)
        Re, te, it = lm_solve(res, kind, R.copy(), t.copy())
        lm_iters.append(it)
        lg = se3_log(R.T @ Re, R.T @ (te - t))                                          # log(cur^-1 * est)
        mse = float(lg @ lg)
        R, t = Re, te
        if mse < mse_stop or count > cap:
            if not semantic:
                count += 1
            break
        if not semantic:
            count += 1
    return Rt_to_pose7(R, t), (count if not semantic else count), lm_iters


def align_gicp(sxyz, txyz, init7, k=20, eps=1e-3):
    scov, _ = covariances(sxyz, k, eps)
    tcov, _ = covariances(txyz, k, eps)

    def build(R, t):
        q = transform_points(R, t, sxyz)
        idx, d2 = knn(txyz, q, 1)
        return [(sxyz[i].astype(np.float64), txyz[idx[i, 0]].astype(np.float64), scov[i], tcov[idx[i, 0]], 1.0)
                for i in range(len(sxyz)) if idx[i, 0] >= 0 and float(d2[i, 0]) < 250]

    return _outer(build, 0, init7, 1e-5, 50)


def align_em(sxyz, slab, txyz, tlab, cm, init7, k=20, eps=1e-3):
    N = cm.shape[0]
    scov, sdist = covariances(sxyz, k, eps, slab, N)
    tcov, tdist = covariances(txyz, k, eps, tlab, N)

    def build(R, t):
        q = transform_points(R, t, sxyz)
        idx, d2 = knn(txyz, q, 4)
        out = []
        for i in range(len(sxyz)):
            for c in range(4):
                j = idx[i, c]
                if j < 0 or not float(d2[i, c]) < 250:
                    continue
                w = sum((tdist[j] @ cm[:, s]) * (sdist[i] @ cm[:, s]) for s in range(N))  # em_icp.hpp:84-89 as written
                ps, pt = sxyz[i].astype(np.float64), txyz[j].astype(np.float64)
                if not probability_is_nonzero(R, t, ps, pt, scov[i], tcov[j]):
                    w = 0.0
                out.append((ps, pt, scov[i], tcov[j], w))
        return out

    return _outer(build, 2, init7, 1e-5, 50)


def align_semantic(sxyz, slab, txyz, tlab, init7, k=20, eps=1e-3):
    def split(lab):
        order = []
        for l in lab:
            if l not in order:
                order.append(l)
        return order

    s_classes = split(list(slab))
    s_idx = {l: np.nonzero(slab == l)[0] for l in s_classes}
    t_idx = {l: np.nonzero(tlab == l)[0] for l in split(list(tlab))}
    scov = {l: covariances(sxyz[s_idx[l]], k, eps)[0] for l in s_classes}               # neighbours within the class
    tcov = {l: covariances(txyz[t_idx[l]], k, eps)[0] for l in t_idx}

    def build(R, t):
        out = []
        for l in s_classes:
            if l not in t_idx or not len(s_idx[l]) > 400:                                # semantic_icp.hpp:50-51
                continue
            sp, tp = sxyz[s_idx[l]], txyz[t_idx[l]]
            idx, d2 = knn(tp, transform_points(R, t, sp), 1)
            for i in range(len(sp)):
                if idx[i, 0] >= 0 and float(d2[i, 0]) < 250:
                    out.append((sp[i].astype(np.float64), tp[idx[i, 0]].astype(np.float64), scov[l][i], tcov[l][idx[i, 0]], 1.0))
        return out

    return _outer(build, 1, init7, 1e-3, 35, semantic=True)


# ---------------------------------------------------------------- getFusedLabels (impl/em_icp.hpp:202-268)
def fused_labels(sxyz, slab, txyz, tlab, cm, pose7, k=20, eps=1e-3):
    N = cm.shape[0]
    scov, sdist = covariances(sxyz, k, eps, slab, N)
    tcov, tdist = covariances(txyz, k, eps, tlab, N)
    R, t = pose7_to_Rt(np.asarray(pose7, dtype=np.float64))
    idx, d2 = knn(txyz, transform_points(R, t, sxyz), 4)
    out = np.zeros(len(sxyz), dtype=np.uint32)
    margin = np.zeros(len(sxyz))
    for i in range(len(sxyz)):
        sprob = np.zeros(N)
        for c in range(4):
            j = idx[i, c]
            if j < 0 or not float(d2[i, c]) < 250:
                continue
            gate = 1.0 if probability_is_nonzero(R, t, sxyz[i].astype(np.float64), txyz[j].astype(np.float64), scov[i], tcov[j]) else 0.0
            for s in range(N):
                sprob[s] += (tdist[j] @ cm[:, s]) * (sdist[i] @ cm[:, s]) * gate
        best, best_s = 0.0, 0
        for s in range(N):                                                              # first strict maximum (:256-262)
            if sprob[s] > best:
                best, best_s = sprob[s], s
        out[i] = best_s + 1
        srt = np.sort(sprob)
        margin[i] = srt[-1] - srt[-2] if N > 1 else 1.0
    return out, margin
