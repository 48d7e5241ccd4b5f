"""ctypes binding of libsicp_b200.so (include/sicp_b200.h) plus thin Python mirrors of the reference classes.

The reference's host language is C++ (the drop-in facade is semantic-icp_b200/facade/*.h); this module exists so
that tests/, bench.py and __graft_entry__.py can drive the C ABI.  There is NO CPU fallback: if the shared library
is missing, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("SICP_LIB") or os.path.join(PKG_ROOT, "lib", "libsicp_b200.so")  # SICP_LIB: A/B builds in tools/

ALGO_GICP, ALGO_SEMANTIC, ALGO_EM = 0, 1, 2
CLOUD_WHOLE, CLOUD_PER_CLASS = 0, 1
STAGES = ["build", "cov", "knn", "estep", "lm", "s5", "s6", "s7"]


class SicpError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [("k_cov", C.c_int), ("epsilon", C.c_double), ("n_classes", C.c_int), ("confusion", C.c_void_p),
                ("gate_d2", C.c_double), ("min_class_points", C.c_int), ("max_lm_iterations", C.c_int), ("profile", C.c_int),
                ("max_concurrent", C.c_int), ("reserved", C.c_int * 6)]


class Result(C.Structure):
    _fields_ = [("pose7", C.c_double * 7), ("outer_iter", C.c_int), ("lm_iters_total", C.c_int), ("final_cost", C.c_double),
                ("n_corr_last", C.c_int), ("flags", C.c_int), ("lm_evals_total", C.c_int), ("gpu_launches", C.c_int), ("d2h_bytes", C.c_int), ("reserved0", C.c_int), ("lm_cycles", C.c_double * 6),
                ("stage_ms", C.c_float * 8), ("stage_launches", C.c_int * 8), ("pass_pose7", (C.c_double * 7) * 64),
                ("pass_lm_iters", C.c_int * 64)]

    def to_dict(self):
        n = min(self.outer_iter, 64)
        return dict(pose=np.array(self.pose7[:]), outer_iter=self.outer_iter, lm_iters_total=self.lm_iters_total,
                    final_cost=self.final_cost, n_corr_last=self.n_corr_last, flags=self.flags, lm_evals_total=self.lm_evals_total,
                    gpu_launches=self.gpu_launches, d2h_bytes=self.d2h_bytes, lm_cycles=list(self.lm_cycles[:]), stage_ms=dict(zip(STAGES, self.stage_ms[:])),
                    stage_launches=dict(zip(STAGES, self.stage_launches[:])),
                    pass_pose=np.array([list(self.pass_pose7[i]) for i in range(n)]).reshape(n, 7),
                    pass_lm_iters=np.array(self.pass_lm_iters[:n], dtype=np.int32))


_lib = None
EXPORTS = [
    "sicp_last_error", "sicp_version", "sicp_device_count", "sicp_set_stream", "sicp_options_default", "sicp_cloud_create",
    "sicp_cloud_create_device", "sicp_cloud_destroy", "sicp_cloud_size", "sicp_cloud_precompute", "sicp_cloud_get_covariances",
    "sicp_cloud_get_normals", "sicp_cloud_get_label_distributions", "sicp_cloud_get_label_vectors", "sicp_cloud_get_self_neighbours",
    "sicp_cloud_get_classes", "sicp_knn", "sicp_knn_cloud", "sicp_correspondences", "sicp_evaluate", "sicp_register",
    "sicp_register_batch", "sicp_fused_labels", "sicp_cloud_transform_f32", "sicp_launch_count", "sicp_label_agreement",
    "sicp_pose_errors", "sicp_filter_range", "sicp_iterative_mean", "sicp_pose_fusion",
]


def lib():
    """Load the CUDA library; fail loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SicpError(f"{LIB_PATH} is missing: build it with `make -C semantic-icp_b200` (or __graft_entry__.build()); "
                            "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.sicp_last_error.restype = C.c_char_p
        _lib.sicp_version.restype = C.c_char_p
        for name in EXPORTS:
            getattr(_lib, name)  # AttributeError if the ABI drifts from include/sicp_b200.h
    return _lib


def _check(st):
    if st != 0:
        raise SicpError(f"sicp status {st}: {lib().sicp_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    n = C.c_int()
    _check(lib().sicp_device_count(C.byref(n)))
    return n.value


def launch_count():
    f = lib().sicp_launch_count
    f.restype = C.c_uint64
    return int(f())


def set_stream(handle):
    _check(lib().sicp_set_stream(C.c_void_p(handle)))


def default_options(algo, cm=None, k_cov=20, epsilon=1e-3, profile=False):
    o = Options()
    lib().sicp_options_default(C.c_int(algo), C.byref(o))
    o.k_cov, o.epsilon, o.profile = k_cov, epsilon, int(profile)
    if cm is not None:
        cm = np.ascontiguousarray(cm, dtype=np.float64)
        o._cm_keepalive = cm
        o.n_classes = cm.shape[0]
        o.confusion = cm.ctypes.data
    return o


class Cloud:
    """Device-resident cloud (opaque sicp_cloud handle)."""

    def __init__(self, xyz, labels=None, layout=CLOUD_WHOLE, device=0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        lab = np.ascontiguousarray(labels, dtype=np.uint32) if labels is not None else None
        self.n = xyz.shape[0]
        self._N, self._k, self._eps = 0, 20, 1e-3
        self.h = C.c_void_p()
        _check(lib().sicp_cloud_create(_p(xyz), C.c_size_t(12), _p(lab), C.c_size_t(4), C.c_size_t(self.n), C.c_int(layout),
                                       C.c_int(device), C.byref(self.h)))

    @classmethod
    def from_device(cls, d_xyz_ptr, d_labels_ptr, n, layout=CLOUD_WHOLE, device=0):
        self = cls.__new__(cls)
        self.n = n
        self._N, self._k, self._eps = 0, 20, 1e-3
        self.h = C.c_void_p()
        _check(lib().sicp_cloud_create_device(C.c_void_p(d_xyz_ptr), C.c_void_p(d_labels_ptr) if d_labels_ptr else None, C.c_size_t(n),
                                              C.c_int(layout), C.c_int(device), C.byref(self.h)))
        return self

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            lib().sicp_cloud_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def precompute(self, k_cov=20, epsilon=1e-3, cm=None):
        N = 0 if cm is None else cm.shape[0]
        cmc = np.ascontiguousarray(cm, dtype=np.float64) if cm is not None else None
        _check(lib().sicp_cloud_precompute(self.h, C.c_int(k_cov), C.c_double(epsilon), C.c_int(N), _p(cmc)))
        self._N, self._k, self._eps = N, k_cov, epsilon

    def normals(self):
        out = np.empty((self.n, 3))
        _check(lib().sicp_cloud_get_normals(self.h, _p(out)))
        return out

    def covariances(self):
        out = np.empty((self.n, 3, 3))
        _check(lib().sicp_cloud_get_covariances(self.h, _p(out)))
        return out

    def label_vectors(self):
        out = np.empty((self.n, self._N))
        _check(lib().sicp_cloud_get_label_vectors(self.h, _p(out)))
        return out

    def label_distributions(self):
        out = np.zeros((self.n, self._N))
        _check(lib().sicp_cloud_get_label_distributions(self.h, _p(out)))
        return out

    def self_neighbours(self):
        out = np.empty((self.n, self._k), dtype=np.int32)
        _check(lib().sicp_cloud_get_self_neighbours(self.h, _p(out)))
        return out

    def classes(self):
        n = C.c_int(128)
        labs = np.empty(128, dtype=np.uint32)
        sizes = np.empty(128, dtype=np.int32)
        _check(lib().sicp_cloud_get_classes(self.h, _p(labs), _p(sizes), C.byref(n)))
        return labs[: n.value].copy(), sizes[: n.value].copy()

    def transform_f32(self, pose7):
        out = np.empty((self.n, 3), dtype=np.float32)
        p = np.ascontiguousarray(pose7, dtype=np.float64)
        _check(lib().sicp_cloud_transform_f32(self.h, _p(p), _p(out), C.c_size_t(12)))
        return out


def knn(target: Cloud, queries, k, pose7=None, q_labels=None):
    q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 3)
    nq = q.shape[0]
    idx = np.full((nq, k), -1, dtype=np.int32)
    d2 = np.full((nq, k), np.inf, dtype=np.float32)
    p = np.ascontiguousarray(pose7, dtype=np.float64) if pose7 is not None else None
    ql = np.ascontiguousarray(q_labels, dtype=np.uint32) if q_labels is not None else None
    _check(lib().sicp_knn(target.h, _p(q), _p(ql), C.c_size_t(nq), _p(p), C.c_int(k), _p(idx), _p(d2)))
    return idx, d2


def knn_cloud(target: Cloud, queries: Cloud, k, d_idx_ptr, d_d2_ptr, pose7=None):
    """Device-resident kNN (sicp_knn_cloud): queries are the points of `queries`; outputs go to the device buffers
    d_idx_ptr (int32 [nq*k]) / d_d2_ptr (float32 [nq*k]) in the queries' original order.  Asynchronous on the current stream."""
    p = np.ascontiguousarray(pose7, dtype=np.float64) if pose7 is not None else None
    _check(lib().sicp_knn_cloud(target.h, queries.h, _p(p), C.c_int(k), C.c_void_p(d_idx_ptr), C.c_void_p(d_d2_ptr)))


def correspondences(algo, src: Cloud, tgt: Cloud, opts: Options, pose7):
    kc = 4 if algo == ALGO_EM else 1
    idx = np.full((src.n, kc), -1, dtype=np.int32)
    w = np.zeros((src.n, kc))
    d2 = np.full((src.n, kc), np.inf, dtype=np.float32)
    p = np.ascontiguousarray(pose7, dtype=np.float64)
    _check(lib().sicp_correspondences(C.c_int(algo), src.h, tgt.h, C.byref(opts), _p(p), _p(idx), _p(w), _p(d2)))
    return idx, w, d2


def evaluate(algo, src: Cloud, tgt: Cloud, opts: Options, corr_pose7, eval_pose7):
    cost = C.c_double()
    g, H = np.empty(6), np.empty((6, 6))
    a = np.ascontiguousarray(corr_pose7, dtype=np.float64)
    b = np.ascontiguousarray(eval_pose7, dtype=np.float64)
    _check(lib().sicp_evaluate(C.c_int(algo), src.h, tgt.h, C.byref(opts), _p(a), _p(b), C.byref(cost), _p(g), _p(H)))
    return cost.value, g, H


def register(algo, src: Cloud, tgt: Cloud, opts: Options, init7):
    res = Result()
    p = np.ascontiguousarray(init7, dtype=np.float64)
    _check(lib().sicp_register(C.c_int(algo), src.h, tgt.h, C.byref(opts), _p(p), C.byref(res)))
    return res.to_dict()


def register_batch(algo, srcs, tgts, opts: Options, inits):
    n = len(srcs)
    arr_s = (C.c_void_p * n)(*[s.h for s in srcs])
    arr_t = (C.c_void_p * n)(*[t.h for t in tgts])
    res = (Result * n)()
    p = np.ascontiguousarray(inits, dtype=np.float64).reshape(n, 7)
    _check(lib().sicp_register_batch(C.c_int(algo), C.c_size_t(n), arr_s, arr_t, C.byref(opts), _p(p), res))
    return [r.to_dict() for r in res]


def fused_labels(src: Cloud, tgt: Cloud, opts: Options, pose7):
    out = np.empty(src.n, dtype=np.uint32)
    p = np.ascontiguousarray(pose7, dtype=np.float64)
    _check(lib().sicp_fused_labels(src.h, tgt.h, C.byref(opts), _p(p), _p(out)))
    return out


def label_agreement(src: Cloud, tgt: Cloud, n_labels, gate_d2=25.0, pose7=None, want_pairs=False):
    """exec/roc_metrics.h:21-41 / exec/nyu_metrics.h:36-84 on the device (sicp_label_agreement)."""
    conf = np.zeros((n_labels, n_labels), dtype=np.int64)
    stats = np.zeros(3)
    pairs = np.empty((src.n, 2), dtype=np.uint32) if want_pairs else None
    p = np.ascontiguousarray(pose7, dtype=np.float64) if pose7 is not None else None
    _check(lib().sicp_label_agreement(src.h, tgt.h, _p(p), C.c_double(gate_d2), C.c_int(n_labels), _p(conf), _p(stats), _p(pairs)))
    return dict(confusion=conf, inliers=stats[0], total=stats[1], dist=stats[2], pairs=pairs)


def pose_errors(gt7s, est7s):
    """exec/kitti_metrics.h:31-37 (sicp_pose_errors): rows of (|log diff|^2, |log_SO3 diff|^2, |t diff|^2)."""
    g = np.ascontiguousarray(gt7s, dtype=np.float64).reshape(-1, 7)
    e = np.ascontiguousarray(est7s, dtype=np.float64).reshape(-1, 7)
    out = np.empty((g.shape[0], 3))
    _check(lib().sicp_pose_errors(C.c_size_t(g.shape[0]), _p(g), _p(e), _p(out)))
    return out


def filter_range(xyz, rng, device=0):
    """exec/filter_range.h:6-18 (sicp_filter_range): indices of the points within `rng` of the origin."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    keep = np.empty(xyz.shape[0], dtype=np.uint32)
    n = C.c_size_t()
    _check(lib().sicp_filter_range(_p(xyz), C.c_size_t(12), C.c_size_t(xyz.shape[0]), C.c_double(rng), C.c_int(device), _p(keep), C.byref(n)))
    return keep[: n.value].copy()


# ------------------------------------------------------------------------------------------------------------------
# Python mirrors of the reference classes (same method names / argument meaning as semantic_icp/*.h)


class GICP:
    """semanticicp::GICP<PointT> (gicp.h:14-132)."""

    def __init__(self, k=20, epsilon=0.001, device=0):
        self.k, self.epsilon, self.device = k, epsilon, device
        self._src = self._tgt = None
        self._res = None

    def setSourceCloud(self, xyz):
        self._src = Cloud(xyz, device=self.device)

    def setTargetCloud(self, xyz):
        self._tgt = Cloud(xyz, device=self.device)

    def align(self, init7=None):
        init7 = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64) if init7 is None else init7
        self._res = register(ALGO_GICP, self._src, self._tgt, default_options(ALGO_GICP, k_cov=self.k, epsilon=self.epsilon), init7)
        return self._src.transform_f32(self._res["pose"])

    def getFinalTransFormation(self):
        return self._res["pose"].copy()

    def getOuterIter(self):
        return self._res["outer_iter"]


class EmIterativeClosestPoint:
    """semanticicp::EmIterativeClosestPoint<N> (em_icp.h:16-122); N is taken from the confusion matrix."""

    def __init__(self, k=20, epsilon=0.001, device=0):
        self.k, self.epsilon, self.device = k, epsilon, device
        self._src = self._tgt = self._cm = None
        self._res = None

    def setSourceCloud(self, xyz, labels):
        self._src = Cloud(xyz, labels, device=self.device)

    def setTargetCloud(self, xyz, labels):
        self._tgt = Cloud(xyz, labels, device=self.device)

    def setConfusionMatrix(self, cm):
        self._cm = np.ascontiguousarray(cm, dtype=np.float64)

    def _opts(self):
        return default_options(ALGO_EM, cm=self._cm, k_cov=self.k, epsilon=self.epsilon)

    def align(self, init7=None):
        init7 = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64) if init7 is None else init7
        self._res = register(ALGO_EM, self._src, self._tgt, self._opts(), init7)
        return self._src.transform_f32(self._res["pose"])

    def getFusedLabels(self, pose7):
        return fused_labels(self._src, self._tgt, self._opts(), pose7)

    def getFinalTransFormation(self):
        return self._res["pose"].copy()

    def getOuterIter(self):
        return self._res["outer_iter"]


class SemanticPointCloud:
    """semanticicp::SemanticPointCloud + pcl_2_semantic (semantic_point_cloud.h:15-63, pcl_2_semantic.h:14-42)."""

    def __init__(self, xyz, labels, k=20, epsilon=0.001, device=0):
        self.cloud = Cloud(xyz, labels, layout=CLOUD_PER_CLASS, device=device)
        self.cloud.precompute(k, epsilon)
        self.semanticLabels, self.sizes = self.cloud.classes()


class SemanticIterativeClosestPoint:
    """semanticicp::SemanticIterativeClosestPoint (semantic_icp.h:14-83)."""

    def __init__(self):
        self._src = self._tgt = None
        self._res = None

    def setInputSource(self, cloud: SemanticPointCloud):
        self._src = cloud

    def setInputTarget(self, cloud: SemanticPointCloud):
        self._tgt = cloud

    def align(self, init7=None):
        init7 = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64) if init7 is None else init7
        c = self._src.cloud
        self._res = register(ALGO_SEMANTIC, c, self._tgt.cloud, default_options(ALGO_SEMANTIC, k_cov=c._k, epsilon=c._eps), init7)
        return c.transform_f32(self._res["pose"])

    def getFinalTransFormation(self):
        return self._res["pose"].copy()


def iterative_mean(poses7, max_iterations=100):
    """SemanticIterativeClosestPoint::iterativeMean (impl/semantic_icp.hpp:169-191): (mean pose7, converged)."""
    p = np.ascontiguousarray(poses7, dtype=np.float64).reshape(-1, 7)
    out = np.zeros(7)
    conv = C.c_int(0)
    _check(lib().sicp_iterative_mean(C.c_size_t(len(p)), _p(p), C.c_int(max_iterations), _p(out), C.byref(conv)))
    return out, bool(conv.value)


def pose_fusion(poses7, covs, init7):
    """SemanticIterativeClosestPoint::poseFusion (impl/semantic_icp.hpp:193-265): (fused pose7, LM iterations)."""
    p = np.ascontiguousarray(poses7, dtype=np.float64).reshape(-1, 7)
    c = np.ascontiguousarray(covs, dtype=np.float64).reshape(-1, 36)
    assert len(c) == len(p)
    out = np.zeros(7)
    it = C.c_int(0)
    _check(lib().sicp_pose_fusion(C.c_size_t(len(p)), _p(p), _p(c), _p(np.ascontiguousarray(init7, dtype=np.float64)), _p(out), C.byref(it)))
    return out, it.value
