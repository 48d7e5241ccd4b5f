/* sicp_b200.h — C ABI of the B200-native Semantic-ICP registration hot path (libsicp_b200.so).
 *
 * The reference (kxhit/semantic-icp) has no FFI: its boundary is the header-only C++ API of
 * semantic_icp/{semantic_point_cloud,pcl_2_semantic,gicp,semantic_icp,em_icp}.h.  Each entry point
 * below names the reference interface it replaces (file:line relative to the reference root).
 * The source-compatible C++ facade over this ABI lives in semantic-icp_b200/facade/ and
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns a sicp_status (0 = OK) and
 * never throws; sicp_last_error() returns a thread-local message.  Poses are 7 doubles
 * [qx,qy,qz,qw,tx,ty,tz] (Sophus::SE3d::data(), gicp_cost_function.h:64-70) mapping the SOURCE
 * frame into the TARGET frame.  Labels are 1..N (em_icp.hpp:301).  Point indices returned to the
 * caller are positions in the caller's original array.  There is no CPU fallback: if no CUDA
 * device is usable every compute call fails with SICP_ERR_CUDA.
 */
#ifndef SICP_B200_H_
#define SICP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int sicp_status;
enum {
  SICP_OK = 0,
  SICP_ERR_INVALID = 1,   /* bad argument (null pointer, label outside 1..N, k out of range ...) */
  SICP_ERR_CUDA = 2,      /* CUDA runtime error or no device */
  SICP_ERR_STATE = 3,     /* call order (e.g. covariances requested before precompute) */
  SICP_ERR_UNSUPPORTED = 4
};

/* algorithm selector: which reference class's align() is reproduced */
enum {
  SICP_ALGO_GICP = 0,     /* semanticicp::GICP<PointT>::align                       impl/gicp.hpp:29-175      */
  SICP_ALGO_SEMANTIC = 1, /* semanticicp::SemanticIterativeClosestPoint::align      impl/semantic_icp.hpp:27-166 */
  SICP_ALGO_EM = 2        /* semanticicp::EmIterativeClosestPoint<N>::align         impl/em_icp.hpp:24-200    */
};

/* cloud layout selector */
enum {
  SICP_CLOUD_WHOLE = 0,    /* one search structure over all points (GICP, EM-ICP)                              */
  SICP_CLOUD_PER_CLASS = 1 /* one search structure per label, first-appearance order (SemanticPointCloud)     */
};

typedef struct sicp_cloud sicp_cloud; /* opaque device-resident cloud: SoA points, Morton-sorted search tree,
                                         normals, label vectors */

typedef struct sicp_options {
  int k_cov;                /* covariance neighbours, 20            gicp.h:34 em_icp.h:42 semantic_point_cloud.h:31 */
  double epsilon;           /* smallest PCA eigenvalue, 1e-3        same                                             */
  int n_classes;            /* EM only: N of EmIterativeClosestPoint<N>                       em_icp.h:16          */
  const double* confusion;  /* EM only: N*N row-major confusion matrix (setConfusionMatrix)   em_icp.h:68          */
  double gate_d2;           /* correspondence gate, 250 m^2         gicp.hpp:70 semantic_icp.hpp:69 em_icp.hpp:65  */
  int min_class_points;     /* SEMANTIC only: class used iff source class size > this, 400    semantic_icp.hpp:51  */
  int max_lm_iterations;    /* 400                                  gicp.hpp:143                                   */
  int profile;              /* 1: record CUDA events per stage into sicp_result.stage_ms                            */
  int max_concurrent;       /* sicp_register_batch: registrations in flight on one GPU (0 = default 8)              */
  int reserved[6];
} sicp_options;

/* stage indices of sicp_result.stage_ms / stage_launches */
enum {
  SICP_STAGE_BUILD = 0,  /* upload + Morton sort + tree build (both clouds)  */
  SICP_STAGE_COV = 1,    /* self-kNN(k_cov) + PCA (+ label vectors)          */
  SICP_STAGE_KNN = 2,    /* transform + kNN(k_c), all passes                 */
  SICP_STAGE_ESTEP = 3,  /* E-step weights, all passes                       */
  SICP_STAGE_LM = 4,     /* M-step: residual/Jacobian/reduction + LM update  */
  SICP_STAGE_COUNT = 8
};

typedef struct sicp_result {
  double pose7[7];      /* getFinalTransFormation()                               gicp.h:98 em_icp.h:82 semantic_icp.h:58 */
  int outer_iter;       /* getOuterIter(): number of outer passes executed       gicp.h:104 em_icp.h:88               */
  int lm_iters_total;   /* LM iterations summed over passes                                                            */
  double final_cost;    /* Ceres-style cost (1/2 sum rho) of the last inner solve                                      */
  int n_corr_last;      /* residual blocks in the last pass                                                            */
  int flags;            /* bit0: outer cap reached                                                                     */
  int lm_evals_total;   /* residual(+Jacobian) sweeps over the correspondence list                                     */
  int gpu_launches;     /* kernels that did work in the outer passes of this registration                              */
  int d2h_bytes;        /* bytes read back to the host (control block) for this registration                          */
  int reserved0;
  double lm_cycles[6];  /* diagnostics, SM cycles inside the LM kernels: block 0 sweeps, block 0 waits, control sections
                           (total, partial reduction, state load, LM step) */
  float stage_ms[SICP_STAGE_COUNT];      /* only when options.profile                                                  */
  int stage_launches[SICP_STAGE_COUNT];
  double pass_pose7[64][7];              /* pose after each outer pass (parity tests)                                  */
  int pass_lm_iters[64];
} sicp_result;

/* ---- library ------------------------------------------------------------------------------------------------ */
const char* sicp_last_error(void);
const char* sicp_version(void);
sicp_status sicp_device_count(int* count);
/* Stream all subsequent calls of this host thread are issued on (a cudaStream_t; NULL = legacy default stream). */
sicp_status sicp_set_stream(void* cuda_stream);
void sicp_options_default(int algo, sicp_options* opts);
/* Kernels launched by this host thread since the library was loaded (cloud builds, precompute, passes). */
uint64_t sicp_launch_count(void);

/* ---- clouds -------------------------------------------------------------------------------------------------
 * sicp_cloud_create replaces GICP::setSourceCloud/setTargetCloud (gicp.h:42-63), Em...::set*Cloud
 * (em_icp.h:50-66) [layout WHOLE] and pcl_2_semantic + SemanticPointCloud::addSemanticCloud's kd-tree build
 * (pcl_2_semantic.h:14-42, impl/semantic_point_cloud.hpp:17-23) [layout PER_CLASS: the first-appearance label
 * partition runs on the device]: it copies the points to the device (pageable host buffers may be reused as soon
 * as the call returns; from page-locked buffers the copy is asynchronous on the calling thread's stream, so they
 * must stay unchanged until a call that consumes the cloud has returned — the reference shares the caller's
 * cloud until align() in the same way) and builds the Z-order-sorted search tree(s).  For layout WHOLE the sort and tree build are
 * deferred to the first call that needs them (a registration, a search, a getter) and run on that call's stream;
 * SICP_EAGER_BUILD=1 builds at once.  xyz points to the first x; consecutive points are xyz_stride bytes apart
 * (12 packed, 16 pcl::PointXYZ, 32 pcl::PointXYZL).  labels may be NULL (GICP).  At most 2^26 points.           */
sicp_status sicp_cloud_create(const void* xyz, size_t xyz_stride, const void* labels, size_t label_stride, size_t n,
                              int layout, int device, sicp_cloud** out);
/* Same, inputs already resident in HBM: d_xyz is packed n*3 float32, d_labels n uint32 (or NULL).
 * For both kinds of cloud the EM label-range check (labels in 1..n_classes) uses the range computed by the device build: a
 * registration reports a violation when it completes, sicp_cloud_precompute reports it at once. */
sicp_status sicp_cloud_create_device(const float* d_xyz, const uint32_t* d_labels, size_t n, int layout, int device,
                                     sicp_cloud** out);
void sicp_cloud_destroy(sicp_cloud* cloud);
sicp_status sicp_cloud_size(const sicp_cloud* cloud, size_t* n);

/* Per-point k-neighbour PCA covariance regularised to (1,1,eps) and, when n_classes > 0, the k-neighbour label
 * distribution pushed through the confusion matrix (a_p = CM^T dist_p).  Replaces GICP::computeCovariances
 * (impl/gicp.hpp:177-239), Em...::ComputeCovariances (impl/em_icp.hpp:270-344) and the covariance loop of
 * addSemanticCloud (impl/semantic_point_cloud.hpp:25-84; neighbours restricted to the point's class when the
 * cloud layout is PER_CLASS).  Idempotent for identical parameters. */
sicp_status sicp_cloud_precompute(sicp_cloud* cloud, int k_cov, double epsilon, int n_classes, const double* confusion);
/* Lazy downloads in the caller's original point order (public maps of SemanticPointCloud, getSourceCovariances). */
sicp_status sicp_cloud_get_covariances(const sicp_cloud* cloud, double* out_n_by_9);
sicp_status sicp_cloud_get_normals(const sicp_cloud* cloud, double* out_n_by_3);
sicp_status sicp_cloud_get_label_distributions(const sicp_cloud* cloud, double* out_n_by_N); /* dist_p (em_icp.hpp:301) */
sicp_status sicp_cloud_get_label_vectors(const sicp_cloud* cloud, double* out_n_by_N);       /* a_p = CM^T dist_p     */
sicp_status sicp_cloud_get_self_neighbours(const sicp_cloud* cloud, int32_t* out_n_by_k);    /* kNN of precompute     */
/* pcl_2_semantic label order (first appearance) and class sizes of a PER_CLASS cloud. */
sicp_status sicp_cloud_get_classes(const sicp_cloud* cloud, uint32_t* labels_out, int32_t* sizes_out, int* n_classes_inout);

/* ---- exact k nearest neighbours -----------------------------------------------------------------------------
 * Replaces pcl::transformPointCloud + KdTreeFLANN::nearestKSearch (gicp.hpp:54-69, em_icp.hpp:46-61,
 * semantic_icp.hpp:52-68).  q_xyz: nq packed float32 host points; if pose7 != NULL each query is first mapped
 * by float(R*p+t) in double arithmetic.  For PER_CLASS targets q_labels selects the class tree (queries whose
 * label is absent get idx -1).  Results: k targets minimising (d2_f32, original index), ascending;
 * idx_out[nq*k] original indices (-1 = none), d2_out[nq*k].  k in {1..32}. */
sicp_status sicp_knn(const sicp_cloud* target, const float* q_xyz, const uint32_t* q_labels, size_t nq, const double* pose7,
                     int k, int32_t* idx_out, float* d2_out);
/* Device-resident variant used for the kNN queries/sec metric: queries are the points of `queries` (in its own
 * original order); outputs are device pointers. */
sicp_status sicp_knn_cloud(const sicp_cloud* target, const sicp_cloud* queries, const double* pose7, int k,
                           int32_t* d_idx_out, float* d_d2_out);

/* ---- pieces of one outer pass, exposed for parity tests ----------------------------------------------------- */
/* Correspondences + weights of one pass at `pose7` (gicp.hpp:66-134 / semantic_icp.hpp:48-134 / em_icp.hpp:57-156):
 * idx_out[ns*kc] (original target index or -1 when gated), w_out[ns*kc] (E-step weight; 0/1 for GICP, SEMANTIC),
 * d2_out[ns*kc].  kc = 4 for EM, 1 otherwise. */
sicp_status sicp_correspondences(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* pose7,
                                 int32_t* idx_out, double* w_out, float* d2_out);
/* Ceres-style evaluation at eval_pose7 of the problem whose correspondences were found at corr_pose7:
 * cost = 1/2 sum rho(r^2), g = J^T r (6), H = J^T J (6x6 row-major), with the loss-corrected 6-dof Jacobian
 * (gicp_cost_function.h:27-73 + local_parameterization_se3.h:30-36 + the loss at each call site). */
sicp_status sicp_evaluate(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* corr_pose7,
                          const double* eval_pose7, double* cost, double* g6, double* H36);

/* ---- registration -------------------------------------------------------------------------------------------
 * One align(): covariance precompute for both clouds (if not cached with the same parameters), then outer passes
 * until the reference's stopping rule.  init7 = initTransform (identity for the one-argument align()).          */
sicp_status sicp_register(int algo, sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* init7,
                          sicp_result* out);
/* n_pairs independent registrations issued concurrently on one GPU; src[i]/tgt[i] may repeat (pose sweeps share
 * clouds).  init7s: n_pairs*7 doubles. */
sicp_status sicp_register_batch(int algo, size_t n_pairs, sicp_cloud* const* src, sicp_cloud* const* tgt, const sicp_options* opts,
                                const double* init7s, sicp_result* out);

/* EmIterativeClosestPoint::getFusedLabels (impl/em_icp.hpp:202-268): labels_out[ns] in source original order. */
sicp_status sicp_fused_labels(sicp_cloud* src, sicp_cloud* tgt, const sicp_options* opts, const double* pose7,
                              uint32_t* labels_out);
/* Final-cloud transform with a float 4x4 (gicp.hpp:166-171, em_icp.hpp:192-197, semantic_point_cloud.hpp:105-111):
 * out_xyz (host, out_stride bytes apart) = float math M4f * p for the cloud's points in original order. */
sicp_status sicp_cloud_transform_f32(const sicp_cloud* cloud, const double* pose7, void* out_xyz, size_t out_stride);

/* ---- evaluation steps either side of the path ----------------------------------------
 * Label agreement (exec/roc_metrics.h:21-41, exec/nyu_metrics.h:36-84): for every source point (mapped by pose7 when
 * it is not NULL; the reference passes the already transformed final cloud) the nearest target point; pairs with
 * d2 < gate_d2 (25 in the reference) contribute (label_source, label_target).
 *   confusion_out [n_labels*n_labels] counts, row = source label (may be NULL when n_labels == 0);
 *   stats3_out    { label matches, gated pairs, sum of sqrt(d2) }  -> accuracy = [0]/[1], mean distance = [2]/[1];
 *   pairs_out     [2*n_source] (label_source, label_target) in source order, 0xffffffff where gated out (nullable). */
sicp_status sicp_label_agreement(const sicp_cloud* src, const sicp_cloud* tgt, const double* pose7, double gate_d2, int n_labels,
                                 int64_t* confusion_out, double* stats3_out, uint32_t* pairs_out);
/* SE(3) error of estimates against ground truth (exec/kitti_metrics.h:31-37): diff = GT * est^-1,
 * err3s[3*i..] = { |log(diff)|^2, |log_SO3(diff)|^2, |translation(diff)|^2 }, one device thread per pose pair. */
sicp_status sicp_pose_errors(size_t n, const double* gt7s, const double* est7s, double* err3s);
/* Range filter of a raw scan (exec/filter_range.h:6-18, used at exec/kitti_eval.cc:124-127): indices (ascending) of the
 * points with x*x+y*y+z*z <= range^2 (float products and sums like the reference), replacing its O(n^2) erase loop. */
sicp_status sicp_filter_range(const void* xyz, size_t xyz_stride, size_t n, double range, int device, uint32_t* keep_idx_out,
                              size_t* n_keep_out);

/* ---- fusing per-class pose estimates (dead code in the reference, kept for completeness) ------------
 * SemanticIterativeClosestPoint::iterativeMean (impl/semantic_icp.hpp:169-191): Karcher mean of n poses on SE(3), started at
 * poses7[0], at most max_iterations steps, stopping when |log(new^-1 * old)|^2 < 0.01.  *converged (nullable) = 0 when the
 * step limit was hit ("Iterative Mean Failed"); out7 is the last iterate either way. */
sicp_status sicp_iterative_mean(size_t n, const double* poses7, int max_iterations, double* out7, int* converged);
/* SemanticIterativeClosestPoint::poseFusion (impl/semantic_icp.hpp:193-265): minimise sum_n Huber_10((e_n^T W_n e_n)^2) with
 * e_n = log(T * pose_n^-1), W_n = cov_n^-1 * (mean det cov)^(1/6), by LM from init7 (50,000 iterations, tolerances
 * 1e-4 * Sophus epsilon).  covs36: n row-major 6x6 matrices (translation block first, Sophus tangent order).  A single
 * pose is returned unchanged. */
sicp_status sicp_pose_fusion(size_t n, const double* poses7, const double* covs36, const double* init7, double* out7,
                             int* lm_iterations);

#ifdef __cplusplus
}
#endif
#endif /* SICP_B200_H_ */
