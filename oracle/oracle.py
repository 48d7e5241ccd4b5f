"""ctypes loader for the CPU oracle (oracle/sicp_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (semantic-icp_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsicp_oracle.so")
_lib = None

LOSS_GICP, LOSS_SEMANTIC, LOSS_EM = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "sicp_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Result(C.Structure):
    _fields_ = [("pose7", C.c_double * 7), ("outer_iter", C.c_int), ("lm_iters_total", C.c_int), ("final_cost", C.c_double),
                ("n_corr_last", C.c_int), ("seconds", C.c_double)]


class Trace(C.Structure):
    _fields_ = [("max_passes", C.c_int), ("pose7", C.c_void_p), ("lm_iters", C.c_void_p), ("n_res", C.c_void_p),
                ("cost", C.c_void_p), ("corr0", C.c_void_p), ("w0", C.c_void_p), ("d20", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def set_search_threads(t):
    """Threads of the neighbour-search loops: 0 = same as the `threads` argument of each call (all-cores baseline),
    1 = serial like the reference (gicp.hpp:66,189; em_icp.hpp:57,288)."""
    lib().orc_set_search_threads(C.c_int(int(t)))


def num_threads():
    return int(lib().orc_num_threads())


def knn(tgt, q, k, brute=False, threads=0):
    tgt, q = _f32(tgt), _f32(q)
    nq = q.shape[0]
    idx = np.empty((nq, k), dtype=np.int32)
    d2 = np.empty((nq, k), dtype=np.float32)
    lib().orc_knn(_p(tgt), C.c_int(tgt.shape[0]), _p(q), C.c_int(nq), C.c_int(k), _p(idx), _p(d2), C.c_int(int(brute)),
                  C.c_int(threads or num_threads()))
    return idx, d2


def knn_search_seconds():
    """Seconds the search loop of the last knn() call took (the kd-tree build excluded)."""
    f = lib().orc_knn_search_seconds
    f.restype = C.c_double
    return float(f())


def transform_points(pose7, xyz):
    xyz = _f32(xyz)
    out = np.empty_like(xyz)
    lib().orc_transform_points(_p(_f64(pose7)), _p(xyz), C.c_int(xyz.shape[0]), _p(out))
    return out


def covariances(xyz, k=20, eps=1e-3, labels=None, N=0, threads=0, want_nn=False):
    xyz = _f32(xyz)
    n = xyz.shape[0]
    cov = np.empty((n, 3, 3))
    normals = np.empty((n, 3))
    dist = np.empty((n, N)) if N > 0 else None
    nn = np.empty((n, k), dtype=np.int32) if want_nn else None
    lab = _u32(labels) if labels is not None else None
    lib().orc_covariances(_p(xyz), _p(lab), C.c_int(n), C.c_int(k), C.c_double(eps), C.c_int(N), _p(cov), _p(dist), _p(normals),
                          _p(nn), C.c_int(threads or num_threads()))
    return dict(cov=cov, normals=normals, dist=dist, nn=nn)


def covariances_per_class(xyz, labels, k=20, eps=1e-3, threads=0):
    xyz, lab = _f32(xyz), _u32(labels)
    n = xyz.shape[0]
    cov = np.empty((n, 3, 3))
    normals = np.empty((n, 3))
    lib().orc_covariances_per_class(_p(xyz), _p(lab), C.c_int(n), C.c_int(k), C.c_double(eps), _p(cov), _p(normals),
                                    C.c_int(threads or num_threads()))
    return dict(cov=cov, normals=normals)


def jacobi_svd(A):
    A = _f64(A)
    U = np.empty((3, 3))
    sv = np.empty(3)
    lib().orc_jacobi_svd(_p(A), _p(U), _p(sv))
    return U, sv


def cost_eval(ps, pt, cs, ct, pose7):
    ps, pt, cs, ct, pose7 = _f32(ps), _f32(pt), _f64(cs), _f64(ct), _f64(pose7)
    r, dens = C.c_double(), C.c_double()
    j7, j6 = np.empty(7), np.empty(6)
    lib().orc_cost_eval(_p(ps), _p(pt), _p(cs), _p(ct), _p(pose7), C.byref(r), _p(j7), _p(j6), C.byref(dens))
    return r.value, j7, j6, dens.value


def loss(kind, w, s):
    rho = np.empty(3)
    lib().orc_loss(C.c_int(kind), C.c_double(w), C.c_double(s), _p(rho))
    return rho


def se3_exp(d6):
    out = np.empty(7)
    lib().orc_se3_exp(_p(_f64(d6)), _p(out))
    return out


def se3_log(p7):
    out = np.empty(6)
    lib().orc_se3_log(_p(_f64(p7)), _p(out))
    return out


def se3_mul(a, b):
    out = np.empty(7)
    lib().orc_se3_mul(_p(_f64(a)), _p(_f64(b)), _p(out))
    return out


def se3_inv(a):
    out = np.empty(7)
    lib().orc_se3_inv(_p(_f64(a)), _p(out))
    return out


def se3_plus(a, d6):
    out = np.empty(7)
    lib().orc_se3_plus(_p(_f64(a)), _p(_f64(d6)), _p(out))
    return out


def se3_matrix(a):
    out = np.empty((4, 4))
    lib().orc_se3_matrix(_p(_f64(a)), _p(out))
    return out


def se3_dx(a):
    out = np.empty((7, 6))
    lib().orc_se3_dx(_p(_f64(a)), _p(out))
    return out


def eval_problem(sxyz, scov, txyz, tcov, s_idx, t_idx, w, loss_kind, pose7, threads=1):
    sxyz, txyz, scov, tcov = _f32(sxyz), _f32(txyz), _f64(scov), _f64(tcov)
    s_idx, t_idx = np.ascontiguousarray(s_idx, dtype=np.int32), np.ascontiguousarray(t_idx, dtype=np.int32)
    w = _f64(w) if w is not None else None
    cost = C.c_double()
    g, H = np.empty(6), np.empty((6, 6))
    lib().orc_eval_problem(_p(sxyz), _p(scov), _p(txyz), _p(tcov), _p(s_idx), _p(t_idx), _p(w), C.c_int(len(s_idx)),
                           C.c_int(loss_kind), _p(_f64(pose7)), C.byref(cost), _p(g), _p(H), C.c_int(threads))
    return cost.value, g, H


def lm_solve(sxyz, scov, txyz, tcov, s_idx, t_idx, w, loss_kind, pose7, threads=0):
    sxyz, txyz, scov, tcov = _f32(sxyz), _f32(txyz), _f64(scov), _f64(tcov)
    s_idx, t_idx = np.ascontiguousarray(s_idx, dtype=np.int32), np.ascontiguousarray(t_idx, dtype=np.int32)
    w = _f64(w) if w is not None else None
    x = _f64(pose7).copy()
    it, term, fc = C.c_int(), C.c_int(), C.c_double()
    lib().orc_lm_solve(_p(sxyz), _p(scov), _p(txyz), _p(tcov), _p(s_idx), _p(t_idx), _p(w), C.c_int(len(s_idx)),
                       C.c_int(loss_kind), _p(x), C.byref(it), C.byref(term), C.byref(fc), C.c_int(threads or num_threads()))
    return x, it.value, term.value, fc.value


class _TraceBufs:
    def __init__(self, ns, kc, max_passes=64):
        self.pose7 = np.zeros((max_passes, 7))
        self.lm_iters = np.zeros(max_passes, dtype=np.int32)
        self.n_res = np.zeros(max_passes, dtype=np.int32)
        self.cost = np.zeros(max_passes)
        self.corr0 = np.full((ns, kc), -1, dtype=np.int32)
        self.w0 = np.zeros((ns, kc))
        self.d20 = np.zeros((ns, kc), dtype=np.float32)
        self.c = Trace(max_passes, _p(self.pose7).value, _p(self.lm_iters).value, _p(self.n_res).value, _p(self.cost).value,
                       _p(self.corr0).value, _p(self.w0).value, _p(self.d20).value)


def _finish(res, tb):
    n = res.outer_iter
    return dict(pose=np.array(res.pose7[:]), outer_iter=n, lm_iters_total=res.lm_iters_total, final_cost=res.final_cost,
                n_corr_last=res.n_corr_last, seconds=res.seconds, pass_pose=tb.pose7[:n].copy(), pass_lm_iters=tb.lm_iters[:n].copy(),
                pass_n_res=tb.n_res[:n].copy(), pass_cost=tb.cost[:n].copy(), corr0=tb.corr0, w0=tb.w0, d20=tb.d20)


def align_gicp(sxyz, txyz, init7, k=20, eps=1e-3, threads=0):
    sxyz, txyz = _f32(sxyz), _f32(txyz)
    res, tb = Result(), _TraceBufs(sxyz.shape[0], 1)
    lib().orc_align_gicp(_p(sxyz), C.c_int(sxyz.shape[0]), _p(txyz), C.c_int(txyz.shape[0]), C.c_int(k), C.c_double(eps),
                         _p(_f64(init7)), C.byref(res), C.byref(tb.c), C.c_int(threads or num_threads()))
    return _finish(res, tb)


def align_em(sxyz, slab, txyz, tlab, cm, init7, k=20, eps=1e-3, threads=0):
    sxyz, txyz, slab, tlab, cm = _f32(sxyz), _f32(txyz), _u32(slab), _u32(tlab), _f64(cm)
    N = cm.shape[0]
    res, tb = Result(), _TraceBufs(sxyz.shape[0], 4)
    lib().orc_align_em(_p(sxyz), _p(slab), C.c_int(sxyz.shape[0]), _p(txyz), _p(tlab), C.c_int(txyz.shape[0]), C.c_int(N), _p(cm),
                       C.c_int(k), C.c_double(eps), _p(_f64(init7)), C.byref(res), C.byref(tb.c), C.c_int(threads or num_threads()))
    return _finish(res, tb)


def align_semantic(sxyz, slab, txyz, tlab, init7, k=20, eps=1e-3, threads=0):
    sxyz, txyz, slab, tlab = _f32(sxyz), _f32(txyz), _u32(slab), _u32(tlab)
    res, tb = Result(), _TraceBufs(sxyz.shape[0], 1)
    lib().orc_align_semantic(_p(sxyz), _p(slab), C.c_int(sxyz.shape[0]), _p(txyz), _p(tlab), C.c_int(txyz.shape[0]), C.c_int(k),
                             C.c_double(eps), _p(_f64(init7)), C.byref(res), C.byref(tb.c), C.c_int(threads or num_threads()))
    return _finish(res, tb)


def fused_labels(sxyz, slab, txyz, tlab, cm, pose7, k=20, eps=1e-3, threads=0):
    sxyz, txyz, slab, tlab, cm = _f32(sxyz), _f32(txyz), _u32(slab), _u32(tlab), _f64(cm)
    out = np.empty(sxyz.shape[0], dtype=np.uint32)
    lib().orc_fused_labels(_p(sxyz), _p(slab), C.c_int(sxyz.shape[0]), _p(txyz), _p(tlab), C.c_int(txyz.shape[0]),
                           C.c_int(cm.shape[0]), _p(cm), C.c_int(k), C.c_double(eps), _p(_f64(pose7)), _p(out),
                           C.c_int(threads or num_threads()))
    return out


def label_split(labels):
    labels = _u32(labels)
    n = labels.shape[0]
    cl = np.empty(n, dtype=np.uint32)
    cs = np.empty(n + 1, dtype=np.int32)
    order = np.empty(n, dtype=np.int32)
    nc = lib().orc_label_split(_p(labels), C.c_int(n), _p(cl), _p(cs), _p(order))
    return cl[:nc].copy(), cs[: nc + 1].copy(), order


# ----------------------------------------------------------------------------------------------------------------------
# Evaluation steps either side of the registration path, restated in numpy on top of the
# oracle's exact kNN / SE(3) routines.


def label_agreement(sxyz, slab, txyz, tlab, n_labels, gate=25.0, pose7=None):
    """exec/roc_metrics.h:21-41 and exec/nyu_metrics.h:36-84: 1-NN of every source point in the target; pairs with
    d2 < 25 give (label_source, label_target), the confusion counts, inlier / total counts and the summed distance."""
    sxyz, txyz = _f32(sxyz), _f32(txyz)
    q = transform_points(pose7, sxyz) if pose7 is not None else sxyz
    idx, d2 = knn(txyz, q, 1)
    slab, tlab = _u32(slab), _u32(tlab)
    keep = (idx[:, 0] >= 0) & (d2[:, 0].astype(np.float64) < gate)                     # nyu_metrics.h:56
    ls, lt = slab[keep], tlab[idx[keep, 0]]
    conf = np.zeros((n_labels, n_labels), dtype=np.int64)
    np.add.at(conf, (ls, lt), 1)                                                        # nyu_metrics.h:59
    dist = 0.0
    for v in np.sqrt(d2[keep, 0]):                                                      # float sqrt, sequential double sum (:64)
        dist += float(v)
    pairs = np.full((sxyz.shape[0], 2), 0xFFFFFFFF, dtype=np.uint32)
    pairs[keep, 0], pairs[keep, 1] = ls, lt
    return dict(confusion=conf, inliers=float(np.sum(ls == lt)), total=float(keep.sum()), dist=dist, pairs=pairs)


def pose_errors(gt7, est7):
    """exec/kitti_metrics.h:31-37: diff = GT * est^-1 -> (|log diff|^2, |log_SO3 diff|^2, |t diff|^2)."""
    d = se3_mul(gt7, se3_inv(est7))
    lg = se3_log(d)
    return np.array([float(lg @ lg), float(lg[3:] @ lg[3:]), float(d[4:] @ d[4:])])


def filter_range(xyz, rng):
    """exec/filter_range.h:6-18: drop points with x*x+y*y+z*z (float arithmetic) > range*range (double)."""
    xyz = _f32(xyz)
    r2 = (xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]) + xyz[:, 2] * xyz[:, 2]
    return np.nonzero(~(r2.astype(np.float64) > rng * rng))[0].astype(np.uint32)


def iterative_mean(poses7, max_iterations=100):
    """SemanticIterativeClosestPoint::iterativeMean (impl/semantic_icp.hpp:169-191): (mean pose7, converged)."""
    p = _f64(np.asarray(poses7).reshape(-1, 7))
    out = np.zeros(7)
    conv = C.c_int(0)
    lib().orc_iterative_mean(_p(p), C.c_int(len(p)), C.c_int(max_iterations), _p(out), C.byref(conv))
    return out, bool(conv.value)


def pose_fusion(poses7, covs, init7):
    """SemanticIterativeClosestPoint::poseFusion (impl/semantic_icp.hpp:193-265): (fused pose7, LM iterations)."""
    p = _f64(np.asarray(poses7).reshape(-1, 7))
    c = _f64(np.asarray(covs).reshape(-1, 36))
    out = np.zeros(7)
    it = C.c_int(0)
    lib().orc_pose_fusion(_p(p), _p(c), C.c_int(len(p)), _p(_f64(init7)), _p(out), C.byref(it))
    return out, it.value
