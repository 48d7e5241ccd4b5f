// knn_cov.cu — kernels K1 (self-kNN + moments + 3x3 Jacobi-SVD PCA), K1b (label vectors) and K2 (transform +
// cross kNN).  THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: every result here must round exactly like the
// reference's scalar x86-64 code (no FMA contraction), see SURVEY.md Appendix A.1-A.5.
//
// Replaces: GICP::computeCovariances (impl/gicp.hpp:177-239), EmIterativeClosestPoint::ComputeCovariances
// (impl/em_icp.hpp:270-344), the covariance loop of SemanticPointCloud::addSemanticCloud
// (impl/semantic_point_cloud.hpp:25-84) and transformPointCloud + nearestKSearch in the outer loops
// (impl/gicp.hpp:54-69, impl/em_icp.hpp:46-61, impl/semantic_icp.hpp:52-68).
#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "knn.cuh"
#include "se3.cuh"
#include "kernels.h"

namespace sicp {

#ifndef SICP_WPB
#define SICP_WPB 4
#endif
constexpr int kWarpsPerBlock = SICP_WPB;
constexpr int kThreads = kWarpsPerBlock * 32;

// Which 32-query packet (leaf of the query cloud) a warp takes.  Expensive packets come in runs of neighbouring leaves
// (a dense region whose queries have no target point nearby), and a block's warps share one SM: handing a block FOUR
// CONSECUTIVE leaves puts a whole run on one SM, which then decides the kernel's duration (ncu: slowest SM busy 2.1x the
// average).  Interleaving — warp w of block b takes leaf w * gridDim + b — spreads a run over neighbouring blocks / SMs,
// but measured SLOWER (k = 4: 656 vs 687 M queries/s, k = 20: 209 vs 242): neighbouring packets share tree nodes and
// leaves through L1, which is worth more than the balance.  Kept as a build option (-DSICP_INTERLEAVE=1), default off.
#ifndef SICP_INTERLEAVE
#define SICP_INTERLEAVE 0
#endif
__device__ __forceinline__ int packet_of_warp(int wib) {
#if SICP_INTERLEAVE
  return wib * (int)gridDim.x + (int)blockIdx.x;
#else
  return (int)blockIdx.x * kWarpsPerBlock + wib;
#endif
}

// ------------------------------------------------------------------ Eigen::JacobiSVD<Matrix3d> restated (U's last column)
struct Rot { double c, s; };
__device__ __forceinline__ void make_jacobi(double x, double y, double z, Rot* r) {
  const double deno = 2.0 * fabs(y);
  if (deno < 2.2250738585072014e-308) { r->c = 1; r->s = 0; return; }
  const double tau = (x - z) / deno;
  const double w = sqrt(tau * tau + 1.0);
  const double t = (tau > 0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double sign_t = t > 0 ? 1.0 : -1.0;
  const double n = 1.0 / sqrt(t * t + 1.0);
  r->s = -sign_t * (y / fabs(y)) * fabs(t) * n;
  r->c = n;
}
// W, U are 3x3 row-major in registers; p, q are compile-time after unrolling
__device__ __forceinline__ void svd3_normal(const double* A, double* normal) {
  const double precision = 2.0 * 2.220446049250313e-16;
  const double tiny = 2.2250738585072014e-308;
  double scale = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) scale = fmax(scale, fabs(A[i]));
  if (!(scale > 0) || !isfinite(scale)) scale = 1.0;
  double W[9], U[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
  for (int i = 0; i < 9; i++) W[i] = A[i] / scale;
  double maxDiag = fmax(fabs(W[0]), fmax(fabs(W[4]), fabs(W[8])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 100) {
    finished = true;
#pragma unroll
    for (int p = 1; p < 3; p++) {
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (q < p) {
          const double threshold = fmax(tiny, precision * maxDiag);
          if (fabs(W[3 * p + q]) > threshold || fabs(W[3 * q + p]) > threshold) {
            finished = false;
            const double m00 = W[3 * p + p], m01 = W[3 * p + q], m10 = W[3 * q + p], m11 = W[3 * q + q];
            Rot rot1;
            const double t = m00 + m11, d = m10 - m01;
            if (fabs(d) < tiny) { rot1.s = 0; rot1.c = 1; }
            else { const double u = t / d; const double tmp = sqrt(1.0 + u * u); rot1.s = 1.0 / tmp; rot1.c = u / tmp; }
            const double n00 = rot1.c * m00 + rot1.s * m10, n01 = rot1.c * m01 + rot1.s * m11;
            const double n11 = -rot1.s * m01 + rot1.c * m11;
            Rot jr;
            make_jacobi(n00, n01, n11, &jr);
            const Rot jrt{jr.c, -jr.s};
            const Rot jl{rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
#pragma unroll
            for (int i = 0; i < 3; i++) {  // W.applyOnTheLeft(p,q,jl)
              const double xi = W[3 * p + i], yi = W[3 * q + i];
              W[3 * p + i] = jl.c * xi + jl.s * yi;
              W[3 * q + i] = -jl.s * xi + jl.c * yi;
            }
            const Rot jlt{jl.c, -jl.s};
#pragma unroll
            for (int i = 0; i < 3; i++) {  // U.applyOnTheRight(p,q,jl^T)
              const double xi = U[3 * i + p], yi = U[3 * i + q];
              U[3 * i + p] = jlt.c * xi - jlt.s * yi;
              U[3 * i + q] = jlt.s * xi + jlt.c * yi;
            }
#pragma unroll
            for (int i = 0; i < 3; i++) {  // W.applyOnTheRight(p,q,jr)
              const double xi = W[3 * i + p], yi = W[3 * i + q];
              W[3 * i + p] = jr.c * xi - jr.s * yi;
              W[3 * i + q] = jr.s * xi + jr.c * yi;
            }
            maxDiag = fmax(maxDiag, fmax(fabs(W[3 * p + p]), fabs(W[3 * q + q])));
          }
        }
      }
    }
  }
  double sv[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double a = fabs(W[4 * i]);
    sv[i] = a;
    if (a != 0) { const double s = W[4 * i] / a; U[i] *= s; U[3 + i] *= s; U[6 + i] *= s; }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) sv[i] *= scale;
  // the column that ends up last after Eigen's descending selection sort
  int c1 = 1, c2 = 2;  // columns in places 1 and 2 (place 0 is not needed)
  {  // i = 0: first maximum of (sv0, sv1, sv2)
    int pos = 0; double mx = sv[0];
    if (sv[1] > mx) { mx = sv[1]; pos = 1; }
    if (sv[2] > mx) { mx = sv[2]; pos = 2; }
    if (mx != 0 && pos != 0) { const double ts = sv[0]; sv[0] = sv[pos]; sv[pos] = ts; if (pos == 1) c1 = 0; else c2 = 0; }
    if (mx != 0) {  // i = 1
      if (sv[2] > sv[1] && sv[2] != 0) { const int tc = c1; c1 = c2; c2 = tc; }
    }
  }
  (void)c1;
  normal[0] = U[c2]; normal[1] = U[3 + c2]; normal[2] = U[6 + c2];
}

// ------------------------------------------------------------------ K1: self kNN + PCA
// moments in kNN order (impl/gicp.hpp:198-222): float*float products rounded to float, double sums, divisor k
struct Moments {
  double mean[3] = {0, 0, 0};
  double c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
  __device__ __forceinline__ void add(const float4& p) {
    mean[0] += p.x; mean[1] += p.y; mean[2] += p.z;
    c00 += __fmul_rn(p.x, p.x);
    c10 += __fmul_rn(p.y, p.x); c11 += __fmul_rn(p.y, p.y);
    c20 += __fmul_rn(p.z, p.x); c21 += __fmul_rn(p.z, p.y); c22 += __fmul_rn(p.z, p.z);
  }
  __device__ __forceinline__ void normal(int k, double* nv) {
    const double kd = (double)k;
    mean[0] /= kd; mean[1] /= kd; mean[2] /= kd;
    double A[9];
    c00 /= kd; c00 -= mean[0] * mean[0];
    c10 /= kd; c10 -= mean[1] * mean[0];
    c11 /= kd; c11 -= mean[1] * mean[1];
    c20 /= kd; c20 -= mean[2] * mean[0];
    c21 /= kd; c21 -= mean[2] * mean[1];
    c22 /= kd; c22 -= mean[2] * mean[2];
    A[0] = c00; A[1] = c10; A[2] = c20; A[3] = c10; A[4] = c11; A[5] = c21; A[6] = c20; A[7] = c21; A[8] = c22;
    svd3_normal(A, nv);
  }
};

#ifndef SICP_COV_MINB
#define SICP_COV_MINB 1   // resident CTAs per SM the register allocation is held to (1: no cap, 88 registers at k = 20)
#endif
template <int K>
__global__ void __launch_bounds__(kThreads, SICP_COV_MINB) self_knn_pca_kernel(CloudView cv, const int* __restrict__ slot_of_orig, int k, double* __restrict__ nrm,
                                                                int* __restrict__ selfnn, uint8_t* __restrict__ nbr_label) {
  __shared__ WarpScratch s_ws[kWarpsPerBlock];
  __shared__ Segment s_seg[kWarpsPerBlock];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leaf = packet_of_warp(wib);
  if (leaf * kLeaf >= cv.nslots) return;
  const int sid = cv.seg_of_leaf[leaf];
  if (lane == 0) s_seg[wib] = cv.seg[sid];
  __syncwarp();
  const Segment& sg = s_seg[wib];
  const int slot = leaf * kLeaf + lane;
  const float4 me = cv.pts[slot];
  const bool valid = __float_as_int(me.w) >= 0;
  TopK<K> L;
  L.init();
  knn_search<K>(cv, sg, me.x, me.y, me.z, valid, L, s_ws[wib]);
  if (!valid) return;
  Moments mo;
#pragma unroll
  for (int j = 0; j < K; j++) {
    const int nslot = (j < k && L.orig(j) >= 0) ? slot_of_orig[L.orig(j)] : -1;
    if (nslot >= 0) mo.add(cv.pts[nslot]);
    if (j < k) {
      if (selfnn) selfnn[(size_t)slot * k + j] = nslot;
      if (nbr_label) nbr_label[(size_t)slot * kMaxK + j] = nslot >= 0 ? (uint8_t)cv.label[nslot] : 0;
    }
  }
  double nv[3];
  mo.normal(k, nv);
  reinterpret_cast<double4*>(nrm)[slot] = make_double4(nv[0], nv[1], nv[2], 0.0);
}

// ------------------------------------------------------------------ K1b: label distribution -> a_p = CM^T dist_p
// one warp per point; dist[b] is the repeated f64 sum of 1/k (em_icp.hpp:279,301), a[s] = sum_b dist[b]*CM[b][s].
// Lane j holds the label of neighbour j.  Only the DISTINCT labels among the k neighbours (typically 1-3 of N) are
// visited, in ascending label order — the same sequence of non-zero terms as the reference's loop over b = 0..N-1 —
// each found with one warp-wide min-reduction.  dist_out (parity tests) must be zeroed by the caller.
__global__ void label_vector_kernel(CloudView cv, int k, int N, const double* __restrict__ cm, const uint8_t* __restrict__ nbr_label,
                                    double* __restrict__ avec, double* __restrict__ dist_out) {
  const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= cv.nslots) return;
  const bool valid = __float_as_int(cv.pts[slot].w) >= 0;  // padding slots carry no neighbour labels
  unsigned lab = (valid && lane < k) ? nbr_label[(size_t)slot * kMaxK + lane] : 0u;
  if (lab > (unsigned)N) lab = 0u;  // labels were range-checked on the host; keeps the CM row index in bounds regardless
  const double inc = 1.0 / (double)k;
  double a0 = 0, a1 = 0;
  unsigned remaining = __ballot_sync(kFull, lab != 0u);
  while (remaining) {
    const unsigned b = __reduce_min_sync(kFull, ((remaining >> lane) & 1u) ? lab : 0xffffffffu);  // smallest label not yet visited
    const unsigned grp = __ballot_sync(kFull, lab == b);
    const int cnt = __popc(grp);
    double d = 0;
    for (int i = 0; i < cnt; i++) d += inc;
    if (dist_out && lane == 0) dist_out[(size_t)slot * N + (b - 1)] = d;
    const double* row = cm + (size_t)(b - 1) * N;
    if (lane < N) a0 += d * __ldg(&row[lane]);
    if (lane + 32 < N) a1 += d * __ldg(&row[lane + 32]);
    remaining &= ~grp;
  }
  if (lane < N) avec[(size_t)slot * N + lane] = a0;
  if (lane + 32 < N) avec[(size_t)slot * N + lane + 32] = a1;
}

// ------------------------------------------------------------------ K2: transform + cross kNN
// p' = float(R p + t) with f64 left-to-right arithmetic (pcl::transformPointCloud<PointT,double>)
__device__ __forceinline__ void transform_rn(const double* R, const double* t, const float4& p, float* q) {
  const double x = p.x, y = p.y, z = p.z;
#pragma unroll
  for (int r = 0; r < 3; r++)
    q[r] = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(R[3 * r], x), __dmul_rn(R[3 * r + 1], y)), __dmul_rn(R[3 * r + 2], z)), t[r]));
}

template <int K>
__global__ void __launch_bounds__(kThreads) cross_knn_kernel(CloudView sv, CloudView tv, const int* __restrict__ tslot_of_orig,
                                                             const double* __restrict__ pose7,
                                                             const int* __restrict__ stop, const int* __restrict__ tseg_of_sseg, int kc,
                                                             int* __restrict__ corr, float* __restrict__ d2out) {
  if (stop && *stop) return;
  __shared__ WarpScratch s_ws[kWarpsPerBlock];
  __shared__ Segment s_seg[kWarpsPerBlock];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leaf = packet_of_warp(wib);
  if (leaf * kLeaf >= sv.nslots) return;
  const int slot = leaf * kLeaf + lane;
  const int tsid = tseg_of_sseg ? tseg_of_sseg[sv.seg_of_leaf[leaf]] : 0;
  if (tsid < 0) {  // class absent from the target or too small (semantic_icp.hpp:50-51)
    for (int c = 0; c < kc; c++) { corr[(size_t)slot * kc + c] = -1; d2out[(size_t)slot * kc + c] = INFINITY; }
    return;
  }
  if (lane == 0) s_seg[wib] = tv.seg[tsid];
  __syncwarp();
  const Segment& sg = s_seg[wib];
  const float4 me = sv.pts[slot];
  const bool valid = __float_as_int(me.w) >= 0;
  float q[3] = {me.x, me.y, me.z};
  if (pose7) {
    double R[9], t[3] = {pose7[4], pose7[5], pose7[6]};
    quat_to_R(pose7, R);
    transform_rn(R, t, me, q);
  }
  TopK<K> L;
  L.init();
  knn_search<K>(tv, sg, q[0], q[1], q[2], valid, L, s_ws[wib]);
#pragma unroll
  for (int c = 0; c < K; c++)
    if (c < kc) {
      const int o = valid ? L.orig(c) : -1;
      corr[(size_t)slot * kc + c] = o >= 0 ? tslot_of_orig[o] : -1;
      d2out[(size_t)slot * kc + c] = o >= 0 ? L.dist(c) : INFINITY;
    }
}

// (query slot order, target slots) -> (query original order, target original indices)
__global__ void unsort_knn_kernel(CloudView qv, CloudView tv, int k, const int* __restrict__ corr, const float* __restrict__ d2,
                                  int32_t* __restrict__ idx_out, float* __restrict__ d2_out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= qv.nslots) return;
  const int o = __float_as_int(qv.pts[slot].w);
  if (o < 0) return;
  for (int c = 0; c < k; c++) {
    const int ts = corr[(size_t)slot * k + c];
    idx_out[(size_t)o * k + c] = ts >= 0 ? __float_as_int(tv.pts[ts].w) : -1;
    d2_out[(size_t)o * k + c] = d2[(size_t)slot * k + c];
  }
}
__global__ void unsort_selfnn_kernel(CloudView cv, int k, const int* __restrict__ selfnn, int32_t* __restrict__ out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= cv.nslots) return;
  const int o = __float_as_int(cv.pts[slot].w);
  if (o < 0) return;
  for (int c = 0; c < k; c++) {
    const int ts = selfnn[(size_t)slot * k + c];
    out[(size_t)o * k + c] = ts >= 0 ? __float_as_int(cv.pts[ts].w) : -1;
  }
}
// slot-ordered SoA/AoS doubles -> original order rows
__global__ void unsort_rows_kernel(CloudView cv, const double* __restrict__ in, int cols, int soa, double* __restrict__ out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= cv.nslots) return;
  const int o = __float_as_int(cv.pts[slot].w);
  if (o < 0) return;
  // soa: 0 = rows of `cols`, 1 = SoA planes, 2 = rows padded to 4 doubles (normals)
  for (int c = 0; c < cols; c++) out[(size_t)o * cols + c] = soa == 1 ? in[(size_t)c * cv.nslots + slot] : soa == 2 ? in[4 * (size_t)slot + c] : in[(size_t)slot * cols + c];
}
// normals (SoA, slot order) -> covariances I - (1-eps) n n^T in original order
__global__ void cov_rows_kernel(CloudView cv, double eps, double* __restrict__ out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= cv.nslots) return;
  const int o = __float_as_int(cv.pts[slot].w);
  if (o < 0) return;
  const double n[3] = {cv.nrm[4 * (size_t)slot], cv.nrm[4 * (size_t)slot + 1], cv.nrm[4 * (size_t)slot + 2]};
  const double kap = 1.0 - eps;
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) out[(size_t)o * 9 + 3 * a + b] = (a == b ? 1.0 : 0.0) - kap * n[a] * n[b];
}

// ------------------------------------------------------------------ host launchers
static int pick_K(int k) { return k <= 1 ? 1 : k <= 4 ? 4 : k <= 20 ? 20 : 32; }

sicp_status launch_self_knn_pca(const sicp_cloud* c, int k, double* d_nrm, int* d_selfnn, uint8_t* d_nbr_label, cudaStream_t st) {
  if (c->nslots == 0) return SICP_OK;
  CloudView cv = c->view();
  const int grid = (c->nleaf + kWarpsPerBlock - 1) / kWarpsPerBlock;
  switch (pick_K(k)) {
    case 1: self_knn_pca_kernel<1><<<grid, kThreads, 0, st>>>(cv, c->d_slot_of_orig, k, d_nrm, d_selfnn, d_nbr_label); break;
    case 4: self_knn_pca_kernel<4><<<grid, kThreads, 0, st>>>(cv, c->d_slot_of_orig, k, d_nrm, d_selfnn, d_nbr_label); break;
    case 20: self_knn_pca_kernel<20><<<grid, kThreads, 0, st>>>(cv, c->d_slot_of_orig, k, d_nrm, d_selfnn, d_nbr_label); break;
    default: self_knn_pca_kernel<32><<<grid, kThreads, 0, st>>>(cv, c->d_slot_of_orig, k, d_nrm, d_selfnn, d_nbr_label); break;
  }
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}

sicp_status launch_cross_knn(const sicp_cloud* src, const sicp_cloud* tgt, const double* d_pose7, const int* d_stop, const int* d_tseg_of_sseg,
                             int kc, int* d_corr, float* d_d2, cudaStream_t st) {
  if (src->nslots == 0) return SICP_OK;
  CloudView sv = src->view(), tv = tgt->view();
  const int grid = (src->nleaf + kWarpsPerBlock - 1) / kWarpsPerBlock;
  switch (pick_K(kc)) {
    case 1: cross_knn_kernel<1><<<grid, kThreads, 0, st>>>(sv, tv, tgt->d_slot_of_orig, d_pose7, d_stop, d_tseg_of_sseg, kc, d_corr, d_d2); break;
    case 4: cross_knn_kernel<4><<<grid, kThreads, 0, st>>>(sv, tv, tgt->d_slot_of_orig, d_pose7, d_stop, d_tseg_of_sseg, kc, d_corr, d_d2); break;
    case 20: cross_knn_kernel<20><<<grid, kThreads, 0, st>>>(sv, tv, tgt->d_slot_of_orig, d_pose7, d_stop, d_tseg_of_sseg, kc, d_corr, d_d2); break;
    default: cross_knn_kernel<32><<<grid, kThreads, 0, st>>>(sv, tv, tgt->d_slot_of_orig, d_pose7, d_stop, d_tseg_of_sseg, kc, d_corr, d_d2); break;
  }
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}

// target segment of every source segment (PER_CLASS clouds), -1 when the class is skipped
sicp_status make_class_map(const sicp_cloud* src, const sicp_cloud* tgt, int min_src_points, int** d_map_out, cudaStream_t st) {
  *d_map_out = nullptr;
  if (tgt->layout != SICP_CLOUD_PER_CLASS) return SICP_OK;
  SICP_REQUIRE(src->layout == SICP_CLOUD_PER_CLASS, "PER_CLASS target needs a PER_CLASS (labelled) source/query cloud");
  std::vector<int> map(std::max(1, src->nseg), -1);
  for (int s = 0; s < src->nseg; s++) {
    if (!(src->h_seg[s].n > min_src_points)) continue;
    for (int t = 0; t < tgt->nseg; t++)
      if (tgt->class_labels[t] == src->class_labels[s]) { map[s] = t; break; }
  }
  SICP_CUDA(cudaMallocAsync(d_map_out, sizeof(int) * map.size(), st));
  SICP_CUDA(cudaMemcpyAsync(*d_map_out, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice, st));
  SICP_CUDA(cudaStreamSynchronize(st));  // `map` is a pageable temporary
  return SICP_OK;
}

}  // namespace sicp

using namespace sicp;

extern "C" {

// debug counters of the stats build (make stats); zeros otherwise
sicp_status sicp_debug_stats(unsigned long long* out8, int reset) {
  for (int i = 0; i < 8; i++) out8[i] = 0;
#ifdef SICP_STATS
  SICP_CUDA(cudaDeviceSynchronize());
  SICP_CUDA(cudaMemcpyFromSymbol(out8, g_stats, sizeof(unsigned long long) * 8));
  if (reset) { unsigned long long z[8] = {0}; SICP_CUDA(cudaMemcpyToSymbol(g_stats, z, sizeof z)); }
#endif
  return SICP_OK;
}

// per-warp traversal trace of the last search kernel (stats build only): out[0..n) scans, out[n..2n) cycles
sicp_status sicp_debug_warp_trace(unsigned* out, int n) {
#ifdef SICP_STATS
  SICP_CUDA(cudaDeviceSynchronize());
  SICP_CUDA(cudaMemcpyFromSymbol(out, g_warp_scans, sizeof(unsigned) * n));
  SICP_CUDA(cudaMemcpyFromSymbol(out + n, g_warp_cycles, sizeof(unsigned) * n));
#else
  for (int i = 0; i < 2 * n; i++) out[i] = 0;
#endif
  return SICP_OK;
}

sicp_status sicp_cloud_precompute(sicp_cloud* c, int k_cov, double eps, int N, const double* cm) {
  return sicp::precompute_cloud(c, k_cov, eps, N, cm, false);
}

}  // extern "C"

// defer_label_check: the label range of a cloud created from DEVICE labels is only known on the device (computed by its
// build).  Fetching it here costs a host synchronisation per cloud on the registration's stream; the batch executor
// instead reads the two words back with the registration's control block and validates them when the registration
// completes (register.cu: Job::check_labels).  The kernels are safe either way (label_vector_kernel ignores labels
// outside 1..N).
sicp_status sicp::precompute_cloud(sicp_cloud* c, int k_cov, double eps, int N, const double* cm, bool defer_label_check) {
  SICP_REQUIRE(c, "cloud is null");
  SICP_REQUIRE(k_cov >= 1 && k_cov <= kMaxK, "k_cov must be in 1..32");
  SICP_REQUIRE(N >= 0 && N <= kMaxClasses, "n_classes must be in 0..64");
  SICP_REQUIRE(N == 0 || cm, "confusion matrix is null");
  SICP_REQUIRE(N == 0 || c->has_labels, "EM precompute needs a labelled cloud");
  // The cache check, the buffers and the ready event are guarded by the cloud's mutex: pairs of an odometry chain share
  // clouds, possibly across host threads.
  std::lock_guard<std::mutex> lock(c->mu);
  if (c->pre_valid && c->pre_k == k_cov && c->pre_eps == eps && c->pre_N == N &&
      (N == 0 || std::memcmp(c->pre_cm.data(), cm, sizeof(double) * N * N) == 0))
    return SICP_OK;
  cudaStream_t st = current_stream();
  SICP_CUDA(cudaSetDevice(c->device));
  SICP_CHECK(ensure_built(c, st));
  if (c->d_nrm || c->d_avec) {
    // Re-precompute with other parameters (rare): registrations on other streams may still be reading the old normals /
    // label vectors, so drain the device before they are overwritten or freed.
    SICP_CUDA(cudaDeviceSynchronize());
  }
  c->pre_valid = false;
  if (!c->d_nrm) SICP_CUDA(cudaMallocAsync(&c->d_nrm, sizeof(double) * 4 * std::max(1, c->nslots), st));
  if (c->d_avec) { SICP_CUDA(cudaFreeAsync(c->d_avec, st)); c->d_avec = nullptr; }
  uint8_t* d_nbr = nullptr;
  double* d_cm = nullptr;
  if (N > 0) {
    if (!c->label_range_known && c->nslots && !defer_label_check) {  // device-resident labels: fetch the range computed at build time
      unsigned h_mm[2];
      SICP_CUDA(cudaMemcpyAsync(h_mm, c->d_bb + 6, 8, cudaMemcpyDeviceToHost, st));
      SICP_CUDA(cudaStreamSynchronize(st));
      c->min_label = h_mm[0]; c->max_label = h_mm[1]; c->label_range_known = true;
    }
    if (c->label_range_known) {
      SICP_REQUIRE(c->nslots == 0 || c->min_label >= 1, "label 0 found: EM-ICP labels must be 1..N");  // em_icp.hpp:301 indexes label-1
      SICP_REQUIRE(c->nslots == 0 || (int)c->max_label <= N, "label exceeds n_classes: EM-ICP labels must be 1..N");
    }
    SICP_CUDA(cudaMallocAsync(&d_nbr, (size_t)kMaxK * std::max(1, c->nslots), st));
    SICP_CUDA(cudaMallocAsync(&c->d_avec, sizeof(double) * N * std::max(1, c->nslots), st));
    SICP_CUDA(cudaMallocAsync(&d_cm, sizeof(double) * N * N, st));
    PinnedBlock pb;
    void* stage = pinned_stage(sizeof(double) * N * N, st, &pb);
    if (!stage) { set_error("pinned staging allocation failed"); return SICP_ERR_CUDA; }
    std::memcpy(stage, cm, sizeof(double) * N * N);
    SICP_CUDA(cudaMemcpyAsync(d_cm, stage, sizeof(double) * N * N, cudaMemcpyHostToDevice, st));
    pinned_release(pb, st);
  }
  SICP_CHECK(launch_self_knn_pca(c, k_cov, c->d_nrm, nullptr, d_nbr, st));
  c->pre_k = k_cov; c->pre_eps = eps; c->pre_N = N;
  if (N > 0) {
    c->pre_cm.assign(cm, cm + N * N);
    if (c->nslots) {
      CloudView cv = c->view();
      label_vector_kernel<<<(unsigned)(((size_t)c->nslots * 32 + 255) / 256), 256, 0, st>>>(cv, k_cov, N, d_cm, d_nbr, c->d_avec, nullptr);
      count_launches(1);
      SICP_CUDA(cudaGetLastError());
    }
    SICP_CUDA(cudaFreeAsync(d_nbr, st));
    SICP_CUDA(cudaFreeAsync(d_cm, st));
  }
  if (!c->ready_ev) SICP_CUDA(cudaEventCreateWithFlags(&c->ready_ev, cudaEventDisableTiming));
  SICP_CUDA(cudaEventRecord(c->ready_ev, st));
  c->pre_valid = true;
  return SICP_OK;
}

extern "C" {

static sicp_status download_rows(const sicp_cloud* c, const double* d_in, int cols, int soa, double* out) {
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_ready(c, st));
  double* d_tmp;
  SICP_CUDA(cudaMallocAsync(&d_tmp, sizeof(double) * cols * std::max<size_t>(1, c->n), st));
  if (c->nslots) unsort_rows_kernel<<<(c->nslots + 255) / 256, 256, 0, st>>>(c->view(), d_in, cols, soa, d_tmp);
  SICP_CUDA(cudaMemcpyAsync(out, d_tmp, sizeof(double) * cols * c->n, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaStreamSynchronize(st));
  SICP_CUDA(cudaFreeAsync(d_tmp, st));
  return SICP_OK;
}

sicp_status sicp_cloud_get_normals(const sicp_cloud* c, double* out) {
  SICP_REQUIRE(c && out, "null argument");
  if (!c->pre_valid) { set_error("precompute has not run"); return SICP_ERR_STATE; }
  SICP_CUDA(cudaSetDevice(c->device));
  return download_rows(c, c->d_nrm, 3, 2, out);
}
sicp_status sicp_cloud_get_covariances(const sicp_cloud* c, double* out) {
  SICP_REQUIRE(c && out, "null argument");
  if (!c->pre_valid) { set_error("precompute has not run"); return SICP_ERR_STATE; }
  SICP_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_ready(c, st));
  double* d_tmp;
  SICP_CUDA(cudaMallocAsync(&d_tmp, sizeof(double) * 9 * std::max<size_t>(1, c->n), st));
  if (c->nslots) cov_rows_kernel<<<(c->nslots + 255) / 256, 256, 0, st>>>(c->view(), c->pre_eps, d_tmp);
  SICP_CUDA(cudaMemcpyAsync(out, d_tmp, sizeof(double) * 9 * c->n, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaStreamSynchronize(st));
  SICP_CUDA(cudaFreeAsync(d_tmp, st));
  return SICP_OK;
}
sicp_status sicp_cloud_get_label_vectors(const sicp_cloud* c, double* out) {
  SICP_REQUIRE(c && out, "null argument");
  if (!c->pre_valid || c->pre_N == 0) { set_error("EM precompute has not run"); return SICP_ERR_STATE; }
  SICP_CUDA(cudaSetDevice(c->device));
  return download_rows(c, c->d_avec, c->pre_N, 0, out);
}
// the two getters below re-run the neighbour search with the debug outputs enabled (parity tests only)
sicp_status sicp_cloud_get_label_distributions(const sicp_cloud* c, double* out) {
  SICP_REQUIRE(c && out, "null argument");
  if (!c->pre_valid || c->pre_N == 0) { set_error("EM precompute has not run"); return SICP_ERR_STATE; }
  SICP_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_ready(c, st));
  const int N = c->pre_N;
  uint8_t* d_nbr; double *d_cm, *d_dist, *d_a, *d_n;
  SICP_CUDA(cudaMallocAsync(&d_nbr, (size_t)kMaxK * std::max(1, c->nslots), st));
  SICP_CUDA(cudaMallocAsync(&d_cm, sizeof(double) * N * N, st));
  SICP_CUDA(cudaMallocAsync(&d_dist, sizeof(double) * N * std::max(1, c->nslots), st));
  SICP_CUDA(cudaMallocAsync(&d_a, sizeof(double) * N * std::max(1, c->nslots), st));
  SICP_CUDA(cudaMallocAsync(&d_n, sizeof(double) * 4 * std::max(1, c->nslots), st));
  SICP_CUDA(cudaMemcpyAsync(d_cm, c->pre_cm.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice, st));
  SICP_CUDA(cudaMemsetAsync(d_dist, 0, sizeof(double) * N * std::max(1, c->nslots), st));
  SICP_CHECK(launch_self_knn_pca(c, c->pre_k, d_n, nullptr, d_nbr, st));
  if (c->nslots) label_vector_kernel<<<(unsigned)(((size_t)c->nslots * 32 + 255) / 256), 256, 0, st>>>(c->view(), c->pre_k, N, d_cm, d_nbr, d_a, d_dist);
  sicp_status rc = download_rows(c, d_dist, N, 0, out);
  cudaFreeAsync(d_nbr, st); cudaFreeAsync(d_cm, st); cudaFreeAsync(d_dist, st); cudaFreeAsync(d_a, st); cudaFreeAsync(d_n, st);
  return rc;
}
sicp_status sicp_cloud_get_self_neighbours(const sicp_cloud* c, int32_t* out) {
  SICP_REQUIRE(c && out, "null argument");
  if (!c->pre_valid) { set_error("precompute has not run"); return SICP_ERR_STATE; }
  SICP_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_ready(c, st));
  const int k = c->pre_k;
  int* d_nn; int32_t* d_out; double* d_n;
  SICP_CUDA(cudaMallocAsync(&d_nn, sizeof(int) * k * std::max(1, c->nslots), st));
  SICP_CUDA(cudaMallocAsync(&d_out, sizeof(int32_t) * k * std::max<size_t>(1, c->n), st));
  SICP_CUDA(cudaMallocAsync(&d_n, sizeof(double) * 4 * std::max(1, c->nslots), st));
  SICP_CHECK(launch_self_knn_pca(c, k, d_n, d_nn, nullptr, st));
  if (c->nslots) unsort_selfnn_kernel<<<(c->nslots + 255) / 256, 256, 0, st>>>(c->view(), k, d_nn, d_out);
  SICP_CUDA(cudaMemcpyAsync(out, d_out, sizeof(int32_t) * k * c->n, cudaMemcpyDeviceToHost, st));
  SICP_CUDA(cudaStreamSynchronize(st));
  cudaFreeAsync(d_nn, st); cudaFreeAsync(d_out, st); cudaFreeAsync(d_n, st);
  return SICP_OK;
}

sicp_status sicp_knn_cloud(const sicp_cloud* tgt, const sicp_cloud* q, const double* pose7, int k, int32_t* d_idx_out, float* d_d2_out) {
  SICP_REQUIRE(tgt && q && d_idx_out && d_d2_out, "null argument");
  SICP_REQUIRE(k >= 1 && k <= kMaxK, "k must be in 1..32");
  SICP_REQUIRE(tgt->device == q->device, "clouds live on different devices");
  SICP_CHECK(validate_pose7(pose7, "sicp_knn_cloud", true));
  SICP_CUDA(cudaSetDevice(tgt->device));
  cudaStream_t st = current_stream();
  SICP_CHECK(ensure_built(tgt, st));
  SICP_CHECK(ensure_built(q, st));
  int* d_map = nullptr;
  SICP_CHECK(make_class_map(q, tgt, -1, &d_map, st));
  double* d_pose = nullptr;
  if (pose7) {  // staged through pinned memory: no host synchronisation on this path
    SICP_CUDA(cudaMallocAsync(&d_pose, 56, st));
    PinnedBlock pb;
    void* stage = pinned_stage(56, st, &pb);
    if (!stage) { set_error("pinned staging allocation failed"); return SICP_ERR_CUDA; }
    std::memcpy(stage, pose7, 56);
    SICP_CUDA(cudaMemcpyAsync(d_pose, stage, 56, cudaMemcpyHostToDevice, st));
    pinned_release(pb, st);
  }
  int* d_corr; float* d_d2;
  SICP_CUDA(cudaMallocAsync(&d_corr, sizeof(int) * k * std::max(1, q->nslots), st));
  SICP_CUDA(cudaMallocAsync(&d_d2, sizeof(float) * k * std::max(1, q->nslots), st));
  SICP_CHECK(launch_cross_knn(q, tgt, d_pose, nullptr, d_map, k, d_corr, d_d2, st));
  if (q->nslots) unsort_knn_kernel<<<(q->nslots + 255) / 256, 256, 0, st>>>(q->view(), tgt->view(), k, d_corr, d_d2, d_idx_out, d_d2_out);
  SICP_CUDA(cudaGetLastError());
  SICP_CUDA(cudaFreeAsync(d_corr, st)); SICP_CUDA(cudaFreeAsync(d_d2, st));
  if (d_pose) SICP_CUDA(cudaFreeAsync(d_pose, st));
  if (d_map) SICP_CUDA(cudaFreeAsync(d_map, st));
  return SICP_OK;
}

sicp_status sicp_knn(const sicp_cloud* tgt, const float* q_xyz, const uint32_t* q_labels, size_t nq, const double* pose7, int k,
                     int32_t* idx_out, float* d2_out) {
  SICP_REQUIRE(tgt && idx_out && d2_out && (q_xyz || nq == 0), "null argument");
  SICP_REQUIRE(k >= 1 && k <= kMaxK, "k must be in 1..32");
  SICP_REQUIRE(tgt->layout == SICP_CLOUD_WHOLE || q_labels, "PER_CLASS target needs query labels");
  SICP_CHECK(validate_pose7(pose7, "sicp_knn", true));
  if (nq == 0) return SICP_OK;
  SICP_CUDA(cudaSetDevice(tgt->device));
  cudaStream_t st = current_stream();
  sicp_cloud* qc = nullptr;
  const int layout = tgt->layout;
  SICP_CHECK(sicp_cloud_create(q_xyz, 12, layout == SICP_CLOUD_PER_CLASS ? q_labels : nullptr, 4, nq, layout, tgt->device, &qc));
  int32_t* d_idx; float* d_d2;
  sicp_status rc = SICP_OK;
  if (cudaMallocAsync(&d_idx, sizeof(int32_t) * k * nq, st) != cudaSuccess || cudaMallocAsync(&d_d2, sizeof(float) * k * nq, st) != cudaSuccess) {
    set_error("allocation failed"); sicp_cloud_destroy(qc); return SICP_ERR_CUDA;
  }
  rc = sicp_knn_cloud(tgt, qc, pose7, k, d_idx, d_d2);
  if (rc == SICP_OK) {
    if (cudaMemcpyAsync(idx_out, d_idx, sizeof(int32_t) * k * nq, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(d2_out, d_d2, sizeof(float) * k * nq, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { set_error("download failed"); rc = SICP_ERR_CUDA; }
  }
  cudaFreeAsync(d_idx, st); cudaFreeAsync(d_d2, st);
  sicp_cloud_destroy(qc);
  return rc;
}

}  // extern "C"
