// metrics.cu — the evaluation steps either side of the registration path, on the device:
//   * label agreement through 1-NN with the d2 < 25 gate      exec/roc_metrics.h:21-41, exec/nyu_metrics.h:36-84
//   * SE(3) error of an estimate against ground truth          exec/kitti_metrics.h:31-37
//   * range filter of a raw scan                               exec/filter_range.h:6-18, exec/kitti_eval.cc:124-127
// They reuse the exact kNN of knn_cov.cu (K2) and the SE(3) arithmetic of se3.cuh.
#include <cub/device/device_select.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "kernels.h"
#include "se3.cuh"

namespace sicp {

constexpr int kMetricThreads = 256;

// One thread per source slot: (label_s, label_t) of every gated 1-NN pair, the n_labels x n_labels confusion counts
// (integer atomics: exact and order-independent) and per-block partial sums of inliers / total / sqrt(d2) (summed in
// block order by the host, so the result is run-to-run deterministic).
__global__ void label_agreement_kernel(CloudView sv, CloudView tv, const int* __restrict__ corr, const float* __restrict__ d2, float gate_d2,
                                       int n_labels, unsigned long long* __restrict__ confusion, double* __restrict__ block_stats,
                                       uint32_t* __restrict__ pairs_out, int* __restrict__ bad_label) {
  __shared__ double s_red[kMetricThreads / 32][3];
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  double inl = 0, tot = 0, dist = 0;
  if (slot < sv.nslots) {
    const int o = __float_as_int(sv.pts[slot].w);
    if (o >= 0) {
      const int ts = corr[slot];
      uint32_t ls = 0xffffffffu, lt = 0xffffffffu;
      if (ts >= 0 && d2[slot] < gate_d2) {  // `nn_dist_sq[0] < 25.0` (roc_metrics.h:34): float compared in double, same truth value
        ls = sv.label[slot];
        lt = tv.label[ts];
        if (ls < (uint32_t)n_labels && lt < (uint32_t)n_labels) atomicAdd(&confusion[(size_t)ls * n_labels + lt], 1ull);
        else if (n_labels > 0) *bad_label = 1;  // Eigen would write out of bounds (nyu_metrics.h:59); reported as an error here
        tot = 1;
        dist = (double)sqrtf(d2[slot]);  // `dist += sqrt(nn_dist_sq[0])`: float sqrt, double accumulation (nyu_metrics.h:64)
        inl = ls == lt ? 1 : 0;
      }
      if (pairs_out) { pairs_out[2 * (size_t)o] = ls; pairs_out[2 * (size_t)o + 1] = lt; }
    }
  }
  for (int off = 16; off; off >>= 1) {
    inl += __shfl_xor_sync(0xffffffffu, inl, off);
    tot += __shfl_xor_sync(0xffffffffu, tot, off);
    dist += __shfl_xor_sync(0xffffffffu, dist, off);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_red[warp][0] = inl; s_red[warp][1] = tot; s_red[warp][2] = dist; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int w = 0; w < kMetricThreads / 32; w++) s += s_red[w][threadIdx.x];
    block_stats[3 * (size_t)blockIdx.x + threadIdx.x] = s;
  }
}

// one thread per pose pair: diff = GT * est^-1 and its three squared norms (exec/kitti_metrics.h:33-37)
__global__ void pose_error_kernel(const double* __restrict__ gt7s, const double* __restrict__ est7s, size_t n, double* __restrict__ err3s) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Pose d = pose_mul(pose_from7(gt7s + 7 * i), pose_inv(pose_from7(est7s + 7 * i)));
  double lg[6];
  pose_log(d, lg);
  err3s[3 * i] = lg[0] * lg[0] + lg[1] * lg[1] + lg[2] * lg[2] + lg[3] * lg[3] + lg[4] * lg[4] + lg[5] * lg[5];
  err3s[3 * i + 1] = lg[3] * lg[3] + lg[4] * lg[4] + lg[5] * lg[5];
  err3s[3 * i + 2] = d.t[0] * d.t[0] + d.t[1] * d.t[1] + d.t[2] * d.t[2];
}

// exec/filter_range.h:12: keep iff !((x*x + y*y + z*z) > range*range), products and sums in float, comparison in double
__global__ void range_flag_kernel(const float* __restrict__ xyz, size_t n, double range2, uint8_t* __restrict__ flags, uint32_t* __restrict__ iota) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  flags[i] = !((double)r2 > range2);
  iota[i] = (uint32_t)i;
}

}  // namespace sicp

using namespace sicp;

extern "C" {

sicp_status sicp_label_agreement(const sicp_cloud* src, const sicp_cloud* tgt, const double* pose7, double gate_d2, int n_labels,
                                 int64_t* confusion_out, double* stats3_out, uint32_t* pairs_out) {
  SICP_REQUIRE(src && tgt && stats3_out, "null argument");
  SICP_REQUIRE(src->has_labels && tgt->has_labels, "label agreement needs labelled clouds");
  SICP_REQUIRE(src->layout == SICP_CLOUD_WHOLE && tgt->layout == SICP_CLOUD_WHOLE, "label agreement needs WHOLE clouds");
  SICP_REQUIRE(n_labels >= 0 && (n_labels == 0 || confusion_out), "confusion_out is null");
  SICP_REQUIRE(src->device == tgt->device, "clouds live on different devices");
  SICP_CHECK(validate_pose7(pose7, "sicp_label_agreement", true));
  SICP_CUDA(cudaSetDevice(src->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_built(src, st));
  SICP_CHECK(ensure_built(tgt, st));
  stats3_out[0] = stats3_out[1] = stats3_out[2] = 0.0;
  if (n_labels) std::memset(confusion_out, 0, sizeof(int64_t) * (size_t)n_labels * n_labels);
  if (src->nslots == 0) return SICP_OK;
  const int nblk = (src->nslots + kMetricThreads - 1) / kMetricThreads;
  int* d_corr = nullptr; float* d_d2 = nullptr; double* d_pose = nullptr; unsigned long long* d_conf = nullptr; double* d_bs = nullptr;
  uint32_t* d_pairs = nullptr; int* d_bad = nullptr;
  std::vector<double> h_bs(3 * (size_t)nblk);
  int h_bad = 0;
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_corr, sizeof(int) * src->nslots, st));
    SICP_CUDA(cudaMallocAsync(&d_d2, sizeof(float) * src->nslots, st));
    SICP_CUDA(cudaMallocAsync(&d_conf, sizeof(unsigned long long) * std::max<size_t>(1, (size_t)n_labels * n_labels), st));
    SICP_CUDA(cudaMallocAsync(&d_bs, sizeof(double) * 3 * nblk, st));
    SICP_CUDA(cudaMallocAsync(&d_bad, sizeof(int), st));
    SICP_CUDA(cudaMemsetAsync(d_conf, 0, sizeof(unsigned long long) * std::max<size_t>(1, (size_t)n_labels * n_labels), st));
    SICP_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if (pairs_out) {
      SICP_CUDA(cudaMallocAsync(&d_pairs, sizeof(uint32_t) * 2 * std::max<size_t>(1, src->n), st));
    }
    if (pose7) {
      SICP_CUDA(cudaMallocAsync(&d_pose, 56, st));
      SICP_CUDA(cudaMemcpyAsync(d_pose, pose7, 56, cudaMemcpyHostToDevice, st));
    }
    SICP_CHECK(launch_cross_knn(src, tgt, d_pose, nullptr, nullptr, 1, d_corr, d_d2, st));
    label_agreement_kernel<<<nblk, kMetricThreads, 0, st>>>(src->view(), tgt->view(), d_corr, d_d2, (float)gate_d2, n_labels, d_conf, d_bs, d_pairs, d_bad);
    count_launches(1);
    SICP_CUDA(cudaGetLastError());
    SICP_CUDA(cudaMemcpyAsync(h_bs.data(), d_bs, sizeof(double) * 3 * nblk, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n_labels) SICP_CUDA(cudaMemcpyAsync(confusion_out, d_conf, sizeof(int64_t) * (size_t)n_labels * n_labels, cudaMemcpyDeviceToHost, st));
    if (pairs_out) SICP_CUDA(cudaMemcpyAsync(pairs_out, d_pairs, sizeof(uint32_t) * 2 * src->n, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  sicp_status rc = body();
  void* bufs[] = {d_corr, d_d2, d_conf, d_bs, d_bad, d_pairs, d_pose};
  for (void* b : bufs) if (b) cudaFreeAsync(b, st);
  SICP_CHECK(rc);
  if (h_bad) { set_error("label >= n_labels in a gated pair (the reference's confusion matrix would be written out of bounds)"); return SICP_ERR_INVALID; }
  for (int b = 0; b < nblk; b++) for (int i = 0; i < 3; i++) stats3_out[i] += h_bs[3 * (size_t)b + i];  // fixed block order
  return SICP_OK;
}

// exec/kitti_metrics.h:31-37: diff = GT * est^-1;  err3 = { |log(diff)|^2, |log_SO3(diff)|^2, |translation(diff)|^2 }
sicp_status sicp_pose_errors(size_t n, const double* gt7s, const double* est7s, double* err3s) {
  SICP_REQUIRE((gt7s && est7s && err3s) || n == 0, "null argument");
  if (n == 0) return SICP_OK;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { set_error("no CUDA device available (libsicp_b200 has no CPU fallback)"); return SICP_ERR_CUDA; }
  cudaStream_t st = current_stream();
  double *d_gt = nullptr, *d_est = nullptr, *d_err = nullptr;
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_gt, 56 * n, st));
    SICP_CUDA(cudaMallocAsync(&d_est, 56 * n, st));
    SICP_CUDA(cudaMallocAsync(&d_err, 24 * n, st));
    SICP_CUDA(cudaMemcpyAsync(d_gt, gt7s, 56 * n, cudaMemcpyHostToDevice, st));
    SICP_CUDA(cudaMemcpyAsync(d_est, est7s, 56 * n, cudaMemcpyHostToDevice, st));
    pose_error_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_gt, d_est, n, d_err);
    count_launches(1);
    SICP_CUDA(cudaGetLastError());
    SICP_CUDA(cudaMemcpyAsync(err3s, d_err, 24 * n, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  const sicp_status rc = body();
  if (d_gt) cudaFreeAsync(d_gt, st);
  if (d_est) cudaFreeAsync(d_est, st);
  if (d_err) cudaFreeAsync(d_err, st);
  return rc;
}

// 6x6 inverse by Gauss-Jordan with partial pivoting (host; pose covariances are tiny); false when singular
static bool inv6(const double* A, double* out) {
  double M[6][12];
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) { M[r][c] = A[6 * r + c]; M[r][6 + c] = r == c ? 1.0 : 0.0; }
  for (int k = 0; k < 6; k++) {
    int piv = k;
    for (int r = k + 1; r < 6; r++) if (std::fabs(M[r][k]) > std::fabs(M[piv][k])) piv = r;
    if (!(std::fabs(M[piv][k]) > 0)) return false;
    if (piv != k) for (int c = 0; c < 12; c++) std::swap(M[k][c], M[piv][c]);
    const double d = M[k][k];
    for (int c = 0; c < 12; c++) M[k][c] /= d;
    for (int r = 0; r < 6; r++)
      if (r != k) { const double f = M[r][k]; if (f != 0) for (int c = 0; c < 12; c++) M[r][c] -= f * M[k][c]; }
  }
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) out[6 * r + c] = M[r][6 + c];
  return true;
}
static double det6(const double* A) {  // LU with partial pivoting
  double M[6][6];
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) M[r][c] = A[6 * r + c];
  double det = 1.0;
  for (int k = 0; k < 6; k++) {
    int piv = k;
    for (int r = k + 1; r < 6; r++) if (std::fabs(M[r][k]) > std::fabs(M[piv][k])) piv = r;
    if (M[piv][k] == 0) return 0.0;
    if (piv != k) { for (int c = 0; c < 6; c++) std::swap(M[k][c], M[piv][c]); det = -det; }
    det *= M[k][k];
    for (int r = k + 1; r < 6; r++) { const double f = M[r][k] / M[k][k]; for (int c = k; c < 6; c++) M[r][c] -= f * M[k][c]; }
  }
  return det;
}

sicp_status sicp_iterative_mean(size_t n, const double* poses7, int max_iterations, double* out7, int* converged) {
  SICP_REQUIRE(poses7 && out7 && n >= 1, "null argument or empty pose list");
  SICP_REQUIRE(n <= 4096, "at most 4096 poses");
  for (size_t i = 0; i < n; i++) SICP_CHECK(validate_pose7(poses7 + 7 * i, "sicp_iterative_mean"));
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { set_error("no CUDA device available (libsicp_b200 has no CPU fallback)"); return SICP_ERR_CUDA; }
  cudaStream_t st = current_stream();
  double *d_in = nullptr, *d_out = nullptr;
  double h_out[8];
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_in, 56 * n, st));
    SICP_CUDA(cudaMallocAsync(&d_out, 64, st));
    SICP_CUDA(cudaMemcpyAsync(d_in, poses7, 56 * n, cudaMemcpyHostToDevice, st));
    SICP_CHECK(launch_iterative_mean(d_in, (int)n, max_iterations, d_out, reinterpret_cast<int*>(d_out + 7), st));
    SICP_CUDA(cudaMemcpyAsync(h_out, d_out, 64, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  const sicp_status rc = body();
  if (d_in) cudaFreeAsync(d_in, st);
  if (d_out) cudaFreeAsync(d_out, st);
  SICP_CHECK(rc);
  std::memcpy(out7, h_out, 56);
  if (converged) std::memcpy(converged, h_out + 7, sizeof(int));
  return SICP_OK;
}

sicp_status sicp_pose_fusion(size_t n, const double* poses7, const double* covs36, const double* init7, double* out7, int* lm_iterations) {
  SICP_REQUIRE(poses7 && covs36 && init7 && out7 && n >= 1, "null argument or empty pose list");
  SICP_REQUIRE(n <= 4096, "at most 4096 poses");
  for (size_t i = 0; i < n; i++) SICP_CHECK(validate_pose7(poses7 + 7 * i, "sicp_pose_fusion"));
  SICP_CHECK(validate_pose7(init7, "sicp_pose_fusion (init7)"));
  if (lm_iterations) *lm_iterations = 0;
  if (n == 1) { std::memcpy(out7, poses7, 56); return SICP_OK; }  // semantic_icp.hpp:223-224
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { set_error("no CUDA device available (libsicp_b200 has no CPU fallback)"); return SICP_ERR_CUDA; }
  // scale = (mean determinant)^(1/6) (semantic_icp.hpp:232-237: 1 / pow(1/det, 1/6)); W_n = cov_n^-1 * scale; pose_n^-1
  double det = 0;
  for (size_t i = 0; i < n; i++) det += det6(covs36 + 36 * i) / (double)n;
  const double scale = 1.0 / std::pow(1.0 / det, 1.0 / 6.0);
  std::vector<double> h(43 * n);  // [n][7] inverse poses, then [n][36] weights
  for (size_t i = 0; i < n; i++) {
    pose_to7(pose_inv(pose_from7(poses7 + 7 * i)), &h[7 * i]);
    double* W = &h[7 * n + 36 * i];
    SICP_REQUIRE(inv6(covs36 + 36 * i, W), "singular pose covariance");
    for (int k = 0; k < 36; k++) W[k] *= scale;
  }
  cudaStream_t st = current_stream();
  double *d_in = nullptr, *d_out = nullptr;
  double h_out[8];
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_in, sizeof(double) * (43 * n + 7), st));
    SICP_CUDA(cudaMallocAsync(&d_out, 64, st));
    SICP_CUDA(cudaMemcpyAsync(d_in, h.data(), sizeof(double) * 43 * n, cudaMemcpyHostToDevice, st));
    SICP_CUDA(cudaMemcpyAsync(d_in + 43 * n, init7, 56, cudaMemcpyHostToDevice, st));
    SICP_CHECK(launch_pose_fusion(d_in, d_in + 7 * n, (int)n, d_in + 43 * n, 50000, d_out, reinterpret_cast<int*>(d_out + 7), st));  // max_num_iterations, semantic_icp.hpp:257
    SICP_CUDA(cudaMemcpyAsync(h_out, d_out, 64, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  const sicp_status rc = body();
  if (d_in) cudaFreeAsync(d_in, st);
  if (d_out) cudaFreeAsync(d_out, st);
  SICP_CHECK(rc);
  std::memcpy(out7, h_out, 56);
  if (lm_iterations) std::memcpy(lm_iterations, h_out + 7, sizeof(int));
  return SICP_OK;
}

sicp_status sicp_filter_range(const void* xyz, size_t xyz_stride, size_t n, double range, int device, uint32_t* keep_idx_out, size_t* n_keep_out) {
  SICP_REQUIRE(n_keep_out && (n == 0 || (xyz && keep_idx_out)), "null argument");
  SICP_REQUIRE(xyz_stride >= 12, "stride too small");
  SICP_REQUIRE(n < (1u << 30), "too many points");
  *n_keep_out = 0;
  if (n == 0) return SICP_OK;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { set_error("no CUDA device available (libsicp_b200 has no CPU fallback)"); return SICP_ERR_CUDA; }
  SICP_CUDA(cudaSetDevice(device));
  cudaStream_t st = current_stream();
  float* d_xyz = nullptr; uint8_t* d_flag = nullptr; uint32_t *d_iota = nullptr, *d_out = nullptr; int* d_num = nullptr; void* d_tmp = nullptr;
  int h_num = 0;
  auto body = [&]() -> sicp_status {
    SICP_CUDA(cudaMallocAsync(&d_xyz, 12 * n, st));
    if (xyz_stride == 12) SICP_CUDA(cudaMemcpyAsync(d_xyz, xyz, 12 * n, cudaMemcpyHostToDevice, st));
    else SICP_CUDA(cudaMemcpy2DAsync(d_xyz, 12, xyz, xyz_stride, 12, n, cudaMemcpyHostToDevice, st));
    SICP_CUDA(cudaMallocAsync(&d_flag, n, st));
    SICP_CUDA(cudaMallocAsync(&d_iota, 4 * n, st));
    SICP_CUDA(cudaMallocAsync(&d_out, 4 * n, st));
    SICP_CUDA(cudaMallocAsync(&d_num, sizeof(int), st));
    range_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_xyz, n, range * range, d_flag, d_iota);
    size_t tmp_bytes = 0;
    SICP_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, d_iota, d_flag, d_out, d_num, (int)n, st));
    SICP_CUDA(cudaMallocAsync(&d_tmp, std::max<size_t>(tmp_bytes, 16), st));
    SICP_CUDA(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_iota, d_flag, d_out, d_num, (int)n, st));  // stable: order preserved
    count_launches(3);
    SICP_CUDA(cudaMemcpyAsync(&h_num, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    SICP_CUDA(cudaMemcpyAsync(keep_idx_out, d_out, 4 * (size_t)h_num, cudaMemcpyDeviceToHost, st));
    SICP_CUDA(cudaStreamSynchronize(st));
    return SICP_OK;
  };
  sicp_status rc = body();
  void* bufs[] = {d_xyz, d_flag, d_iota, d_out, d_num, d_tmp};
  for (void* b : bufs) if (b) cudaFreeAsync(b, st);
  SICP_CHECK(rc);
  *n_keep_out = (size_t)h_num;
  return SICP_OK;
}

}  // extern "C"
