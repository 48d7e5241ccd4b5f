// kernels.h — host-side launchers shared between the translation units of libsicp_b200.
#pragma once
#include "common.cuh"

namespace sicp {

// Device-resident control block of one registration: the outer ICP loop state lives here so that passes can be
// enqueued back to back without a host round trip (kernels return immediately once `converged` is set).
struct RegCtl {
  double pose[7];        // current transform (source -> target)
  int converged;         // set by the LM kernel's epilogue (impl/gicp.hpp:153-155 etc.)
  int outer;             // passes completed
  int lm_iters_total;
  int lm_evals_total;
  int n_corr_last;
  int term_last;
  int flags;
  int n_corr_pass;
  double final_cost;
  double last_mse;
  double pass_pose[64][7];
  int pass_lm_iters[64];
  long long dbg_cycles[8];  // block 0: sweep compute, grid sync + final reduce, LM control, total (accumulated over passes)
};

struct LMConfig {
  int algo;              // SICP_ALGO_*
  int kc;                // correspondences per source point
  double eps;            // PCA epsilon
  double kappa, hk, k4, aa;  // 1-eps, kappa/2, kappa/4, 1-kappa/2: constant-bank operands of the residual sweep
  int max_iter;          // 400
  double mse_stop;       // 1e-5 (GICP, EM) / 1e-3 (SEMANTIC)
  int outer_cap;         // 50 / 35
  int variant;           // shape of the LM kernel (lm.cu: kLmShapes)
  int ctl_share8;        // sweep share of the LM controller block in eighths of a normal block's share (0..8)
};
constexpr int kLmVariants = 5;
constexpr int kLmMaxGrid = 320;       // block partials reserved per workspace
constexpr int kLmSyncDoubles = 32;    // LMSync lives in front of the partials
constexpr size_t kLmPartialsDoubles = kLmSyncDoubles + (size_t)kLmMaxGrid * 28;

// Residual records of one outer pass: written once by the E-step, streamed ~30 times by the LM sweeps of the pass.
// They are stored in GROUP BLOCKS — one contiguous block per 32 consecutive source slots (one warp-iteration of the
// sweep) — so that a sweep warp fetches everything it needs for a group with ONE bulk (TMA) copy into shared memory:
//   candidate c (kCandBytes each):  w[32] f64 | px[32] py[32] pz[32] f32 | n_t.x[32] n_t.y[32] n_t.z[32] f64
//   then the source side:           sx[32] sy[32] sz[32] f32 | n_s.x[32] n_s.y[32] n_s.z[32] f64
// Every field is a lane-indexed run (lane = slot & 31): E-step stores and sweep loads are coalesced / conflict-free.
// Records without a residual (gated, padding) carry w = 0 and zeroed geometry, which keeps the branch-free sweep finite.
struct Rec {
  static constexpr int kW = 0, kPx = 256, kPy = 384, kPz = 512, kNx = 640, kNy = 896, kNz = 1152, kCandBytes = 1408;
  static constexpr int kSx = 0, kSy = 128, kSz = 256, kSnx = 384, kSny = 640, kSnz = 896, kSrcBytes = 1152;
  static constexpr int group_bytes(int kc) { return kc * kCandBytes + kSrcBytes; }  // 6784 (k_c = 4), 2560 (k_c = 1): multiples of 16
};

sicp_status launch_self_knn_pca(const sicp_cloud* c, int k, double* d_nrm, int* d_selfnn, uint8_t* d_nbr_label, cudaStream_t st);
sicp_status launch_cross_knn(const sicp_cloud* src, const sicp_cloud* tgt, const double* d_pose7, const int* d_stop, const int* d_tseg_of_sseg,
                             int kc, int* d_corr, float* d_d2, cudaStream_t st);
sicp_status make_class_map(const sicp_cloud* src, const sicp_cloud* tgt, int min_src_points, int** d_map_out, cudaStream_t st);

// E-step: gate + label-compatibility weight + probability gate (impl/em_icp.hpp:65-89,108; gicp_cost_function.h:75-87)
sicp_status launch_estep(const sicp_cloud* src, const sicp_cloud* tgt, const LMConfig& cfg, double gate_d2, const double* d_pose7,
                         const int* d_stop, int* d_corr, const float* d_d2, char* d_rec, RegCtl* d_ctl, cudaStream_t st);
// M-step: one inner solve (ceres::Solve at impl/gicp.hpp:149-151) + outer-loop bookkeeping, cooperative kernel
int lm_grid_blocks(int device);
int lm_max_grid(int device, int algo, int variant);
sicp_status precompute_cloud(sicp_cloud* c, int k_cov, double eps, int n_classes, const double* cm, bool defer_label_check);  // knn_cov.cu
// cond_handle != 0: the launch is being captured as the last node of a graph WHILE body; the kernel sets the handle to
// "not converged" so that the graph runs another pass without the host
sicp_status launch_lm(const sicp_cloud* src, const LMConfig& cfg, const char* d_rec, RegCtl* d_ctl,
                      double* d_partials, int grid, cudaStream_t st, unsigned long long cond_handle = 0);
// The inner solves of TWO registrations (same algorithm) in one cooperative launch whose sweeping blocks alternate between
// them, so that each problem's control step overlaps the other's sweep (lm.cu: lm_pair_kernel).  cond_handle as above,
// set to "some registration of the pair has not converged".
sicp_status launch_lm_pair(const sicp_cloud* src0, const LMConfig& cfg0, const char* d_rec0, RegCtl* d_ctl0, double* d_partials0,
                           const sicp_cloud* src1, const LMConfig& cfg1, const char* d_rec1, RegCtl* d_ctl1, double* d_partials1,
                           int grid, cudaStream_t st, unsigned long long cond_handle);
// Single evaluation (cost, g, H) at a given pose, for parity tests
sicp_status launch_evaluate(const sicp_cloud* src, const LMConfig& cfg, const char* d_rec,
                            const double* d_pose7, double* d_out28, double* d_partials, int grid, cudaStream_t st);
// fused labels (impl/em_icp.hpp:202-268)
sicp_status launch_fused_labels(const sicp_cloud* src, const sicp_cloud* tgt, double eps, double gate_d2, const double* d_pose7, const int* d_corr,
                                const float* d_d2, uint32_t* d_labels_out, cudaStream_t st);

// pose averaging / fusion of per-class estimates (impl/semantic_icp.hpp:169-265)
sicp_status launch_iterative_mean(const double* d_poses7, int n, int max_iter, double* d_out7, int* d_converged, cudaStream_t st);
sicp_status launch_pose_fusion(const double* d_pinv7s, const double* d_Ws, int n, const double* d_init7, int max_iter, double* d_out7, int* d_iters,
                               cudaStream_t st);

}  // namespace sicp
