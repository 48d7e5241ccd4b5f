// lm.cu — E-step weights (K3) and the device-side M-step (K4 + K5): analytic SE(3) residual/Jacobian evaluation,
// warp-shuffle reduction of J^T J (21) + J^T r (6) + cost (1), and a Ceres-mirroring Levenberg-Marquardt loop that
// runs entirely inside ONE cooperative kernel per outer pass (no host round trips inside an inner solve).
//
// Replaces: GICPCostFunction::Evaluate/Probability (gicp_cost_function.h:27-87), LocalParameterizationSE3
// (local_parameterization_se3.h:17-36), SQLoss (sqloss.h:11-19) + the Ceres loss compositions at
// impl/gicp.hpp:98-104, impl/em_icp.hpp:109-117, impl/semantic_icp.hpp:96, ceres::Solve (impl/gicp.hpp:138-151,
// impl/em_icp.hpp:162-177, impl/semantic_icp.hpp:136-149) and the outer-loop test (impl/gicp.hpp:153-161,
// impl/em_icp.hpp:179-187, impl/semantic_icp.hpp:47,151-158).
//
// Covariances are never materialised: C = I - (1-eps) n n^T, so
//   C_t + R C_s R^T = 2I - kappa (n_t n_t^T + m m^T),  m = R n_s,  kappa = 1 - eps
// is inverted in closed form (rank-2 Woodbury), and the 6-dof Jacobian of r = d^T M d for T*exp(delta) is
//   J_upsilon = -2 c,   J_omega = 2 c x (p_s + C_s c),   c = R^T M d           (verified vs the 1x7 route).
#include <cooperative_groups.h>
#include <cfloat>
#include "common.cuh"
#include "kernels.h"
#include "se3.cuh"

namespace cg = cooperative_groups;

namespace sicp {

constexpr int kLmThreads = 256;
constexpr int kAcc = 28;  // 21 lower-triangular H + 6 g + cost
constexpr unsigned kFullMask = 0xffffffffu;

struct RT { double R[9]; double t[3]; };

__device__ __forceinline__ void load_point(const CloudView& v, int slot, double* p, double* n) {
  const float4 a = __ldg(&v.pts[slot]);
  p[0] = a.x; p[1] = a.y; p[2] = a.z;
  n[0] = __ldg(&v.nrm[slot]);
  n[1] = __ldg(&v.nrm[(size_t)v.nslots + slot]);
  n[2] = __ldg(&v.nrm[2 * (size_t)v.nslots + slot]);
}

// b = M d with M = (2I - kappa(u u^T + v v^T))^-1
__device__ __forceinline__ void apply_Minv(const double* u, const double* v, const double* d, double kappa, double* b) {
  const double c = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
  const double p = u[0] * d[0] + u[1] * d[1] + u[2] * d[2];
  const double q = v[0] * d[0] + v[1] * d[1] + v[2] * d[2];
  const double a = 1.0 - 0.5 * kappa, be = 0.5 * kappa * c;
  const double idet = 1.0 / ((a - be) * (a + be));
  const double g1 = (a * p + be * q) * idet, g2 = (be * p + a * q) * idet;
  const double k4 = 0.25 * kappa;
#pragma unroll
  for (int i = 0; i < 3; i++) b[i] = 0.5 * d[i] + k4 * (g1 * u[i] + g2 * v[i]);
}
__device__ __forceinline__ double det_S(const double* u, const double* v, double kappa) {
  // det(2I - kappa(uu^T+vv^T)) = 2 * det(2 I2 - kappa W^T W) = 2 * ((2-kappa)^2 - kappa^2 c^2)
  const double c = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
  const double a = 2.0 - kappa, be = kappa * c;
  return 2.0 * (a - be) * (a + be);
}

// Ceres loss compositions: rho(s) and rho'(s) at s = r^2.
__device__ __forceinline__ void loss_eval(int algo, double w, double s, double* rho0, double* rho1) {
  if (algo == SICP_ALGO_SEMANTIC) {  // CauchyLoss(1.5)
    const double b = 2.25, c = 1.0 / 2.25;
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    *rho0 = b * log(sum);
    *rho1 = fmax(DBL_MIN, inv);
    return;
  }
  // ComposedLoss(CauchyLoss(3.0) [scaled by w for EM], SQLoss)
  const double v = s + DBL_EPSILON;
  const double g0 = sqrt(v);
  const double g1 = 1.0 / (2.0 * g0);
  const double b = 9.0, c = 1.0 / 9.0;
  const double sum = 1.0 + g0 * c, inv = 1.0 / sum;
  double f0 = b * log(sum), f1 = fmax(DBL_MIN, inv);
  if (algo == SICP_ALGO_EM) { f0 *= w; f1 *= w; }
  *rho0 = f0;
  *rho1 = f1 * g1;
}

template <bool JAC>
__device__ __forceinline__ void accumulate_residual(const CloudView& sv, const CloudView& tv, int slot, int ts, double w, const RT& P,
                                                    double kappa, int algo, double* acc) {
  double ps[3], ns[3], pt[3], nt[3];
  load_point(sv, slot, ps, ns);
  load_point(tv, ts, pt, nt);
  double m[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    m[i] = P.R[3 * i] * ns[0] + P.R[3 * i + 1] * ns[1] + P.R[3 * i + 2] * ns[2];
    d[i] = pt[i] - (P.R[3 * i] * ps[0] + P.R[3 * i + 1] * ps[1] + P.R[3 * i + 2] * ps[2] + P.t[i]);
  }
  double b[3];
  apply_Minv(nt, m, d, kappa, b);
  const double r = d[0] * b[0] + d[1] * b[1] + d[2] * b[2];
  double rho0, rho1;
  loss_eval(algo, w, r * r, &rho0, &rho1);
  acc[27] += 0.5 * rho0;
  if (JAC) {
    double c[3];
#pragma unroll
    for (int i = 0; i < 3; i++) c[i] = P.R[i] * b[0] + P.R[3 + i] * b[1] + P.R[6 + i] * b[2];  // R^T b
    const double nc = kappa * (ns[0] * c[0] + ns[1] * c[1] + ns[2] * c[2]);
    const double e[3] = {ps[0] + c[0] - nc * ns[0], ps[1] + c[1] - nc * ns[1], ps[2] + c[2] - nc * ns[2]};  // p_s + C_s c
    const double sr = sqrt(rho1);
    const double s2 = 2.0 * sr;
    double J[6];
    J[0] = -s2 * c[0]; J[1] = -s2 * c[1]; J[2] = -s2 * c[2];
    J[3] = s2 * (c[1] * e[2] - c[2] * e[1]);
    J[4] = s2 * (c[2] * e[0] - c[0] * e[2]);
    J[5] = s2 * (c[0] * e[1] - c[1] * e[0]);
    const double rc = sr * r;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
      for (int bb = 0; bb <= a; bb++) acc[k++] += J[a] * J[bb];
      acc[21 + a] += J[a] * rc;
    }
  }
}

// ------------------------------------------------------------------ K3: E-step
__global__ void estep_kernel(CloudView sv, CloudView tv, int algo, int kc, double eps, double gate_d2, const double* __restrict__ pose7,
                             const int* __restrict__ stop, int* __restrict__ corr, const float* __restrict__ d2, double* __restrict__ wout, RegCtl* ctl) {
  if (stop && *stop) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < sv.nslots * kc;
  const int slot = live ? r / kc : 0;
  int ts = live ? corr[r] : -1;
  double w = 0.0;
  if (ts >= 0 && !((double)d2[r] < gate_d2)) { ts = -1; corr[r] = -1; }  // `distSq < 250` (gicp.hpp:70, em_icp.hpp:65)
  if (ts >= 0) {
    w = 1.0;
    if (algo == SICP_ALGO_EM) {
      // label compatibility (em_icp.hpp:84-89) with a_p = CM^T dist_p precomputed per point
      const int N = sv.N;
      const double* as = sv.avec + (size_t)slot * N;
      const double* at = tv.avec + (size_t)ts * N;
      double prob = 0.0;
      for (int s = 0; s < N; s++) prob += __ldg(&at[s]) * __ldg(&as[s]);
      // Probability() -> bool (gicp_cost_function.h:75-87): weight kept iff the density is not exactly 0
      RT P;
      quat_to_R(pose7, P.R);
      P.t[0] = pose7[4]; P.t[1] = pose7[5]; P.t[2] = pose7[6];
      double ps[3], ns[3], pt[3], nt[3], m[3], d[3], b[3];
      load_point(sv, slot, ps, ns);
      load_point(tv, ts, pt, nt);
      for (int i = 0; i < 3; i++) {
        m[i] = P.R[3 * i] * ns[0] + P.R[3 * i + 1] * ns[1] + P.R[3 * i + 2] * ns[2];
        d[i] = pt[i] - (P.R[3 * i] * ps[0] + P.R[3 * i + 1] * ps[1] + P.R[3 * i + 2] * ps[2] + P.t[i]);
      }
      const double kappa = 1.0 - eps;
      apply_Minv(nt, m, d, kappa, b);
      const double mahal = -0.5 * (d[0] * b[0] + d[1] * b[1] + d[2] * b[2]);
      const double two_pi = 6.283185307179586;
      const double det2pi = (two_pi * two_pi * two_pi) * det_S(nt, m, kappa);
      const double density = pow(det2pi, -0.5) * exp(mahal);
      if (density == 0.0) prob *= 0.0;  // NaN stays "true" like the bool conversion
      w = prob;
    }
  }
  if (live) wout[r] = w;
  if (ctl) {  // residual blocks of this pass (diagnostics)
    const int cnt = __popc(__ballot_sync(kFullMask, ts >= 0));
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&ctl->n_corr_pass, cnt);
  }
}

// ------------------------------------------------------------------ K4 + K5: LM
struct LMState {
  double x[7], cand[7];
  double tot[kAcc];
  double H[21], g[6], cost;
  double scale[6], diag[6];
  double radius, decrease_factor, x_norm, gmax, model;
  int reuse_diag, last_successful, invalid, iter, evals, term, done;
};
enum { TERM_NO_CONV = 0, TERM_GRADIENT = 1, TERM_PARAMETER = 2, TERM_FUNCTION = 3, TERM_RADIUS = 4, TERM_FAIL = 5 };

__device__ __forceinline__ int tri(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

__device__ bool chol_solve6(const double* A /*6x6 row-major*/, const double* b, double* x) {
  double L[36];
  for (int i = 0; i < 36; i++) L[i] = 0;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[6 * i + j];
      for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k];
      if (i == j) { if (!(s > 0)) return false; L[6 * i + i] = sqrt(s); }
      else L[6 * i + j] = s / L[6 * j + j];
    }
  double y[6];
  for (int i = 0; i < 6; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[6 * i + k] * y[k]; y[i] = s / L[6 * i + i]; }
  for (int i = 5; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * x[k]; x[i] = s / L[6 * i + i]; }
  return true;
}
__device__ double grad_max_norm(const double* x7, const double* g) {
  double ng[6];
  for (int j = 0; j < 6; j++) ng[j] = -g[j];
  double p7[7];
  pose_to7(pose_plus(pose_from7(x7), ng), p7);
  double m = 0;
  for (int i = 0; i < 7; i++) m = fmax(m, fabs(x7[i] - p7[i]));
  return m;
}
__device__ double norm7(const double* a) { double s = 0; for (int i = 0; i < 7; i++) s += a[i] * a[i]; return sqrt(s); }

// after the evaluation at x0 (iteration zero of ceres::TrustRegionMinimizer)
__device__ void lm_init(LMState& S) {
  for (int i = 0; i < 21; i++) S.H[i] = S.tot[i];
  for (int i = 0; i < 6; i++) S.g[i] = S.tot[21 + i];
  S.cost = S.tot[27];
  for (int j = 0; j < 6; j++) S.scale[j] = 1.0 / (1.0 + sqrt(S.H[tri(j, j)]));  // Jacobi scaling, computed once
  S.gmax = grad_max_norm(S.x, S.g);
  S.x_norm = norm7(S.x);
  S.radius = 1e4; S.decrease_factor = 2.0; S.reuse_diag = 0; S.last_successful = 1; S.invalid = 0; S.iter = 0; S.term = TERM_NO_CONV;
}
// top of the minimizer loop up to the candidate point; sets S.done when the solve terminates
__device__ void lm_propose(LMState& S, int max_iter) {
  const double gtol = 0.1 * kSophusEps, ptol = 1e-8;
  for (;;) {
    if (S.iter >= max_iter) { S.term = TERM_NO_CONV; S.done = 1; return; }
    if (S.last_successful && S.gmax <= gtol) { S.term = TERM_GRADIENT; S.done = 1; return; }
    if (S.radius <= 1e-32) { S.term = TERM_RADIUS; S.done = 1; return; }
    S.iter++;
    double Hs[36], gs[6];
    for (int a = 0; a < 6; a++) {
      gs[a] = S.g[a] * S.scale[a];
      for (int b = 0; b < 6; b++) Hs[6 * a + b] = S.H[tri(a, b)] * S.scale[a] * S.scale[b];
    }
    if (!S.reuse_diag) for (int j = 0; j < 6; j++) S.diag[j] = fmin(fmax(Hs[7 * j], 1e-6), 1e32);
    double A[36];
    for (int i = 0; i < 36; i++) A[i] = Hs[i];
    for (int j = 0; j < 6; j++) A[7 * j] += S.diag[j] / S.radius;
    double y[6], step[6];
    const bool ok = chol_solve6(A, gs, y);
    S.reuse_diag = 1;
    double model = 0;
    if (ok) {
      double sg = 0, sHs = 0;
      for (int a = 0; a < 6; a++) step[a] = -y[a];
      for (int a = 0; a < 6; a++) { sg += step[a] * gs[a]; for (int b = 0; b < 6; b++) sHs += step[a] * Hs[6 * a + b] * step[b]; }
      model = -sg - 0.5 * sHs;
    }
    if (!ok || !(model > 0)) {
      if (++S.invalid >= 5) { S.term = TERM_FAIL; S.done = 1; return; }
      S.radius /= S.decrease_factor; S.decrease_factor *= 2.0; S.last_successful = 0;
      continue;
    }
    S.invalid = 0;
    S.model = model;
    double delta[6];
    for (int j = 0; j < 6; j++) delta[j] = step[j] * S.scale[j];
    pose_to7(pose_plus(pose_from7(S.x), delta), S.cand);
    double sn = 0;
    for (int i = 0; i < 7; i++) sn += (S.x[i] - S.cand[i]) * (S.x[i] - S.cand[i]);
    if (sqrt(sn) <= ptol * (S.x_norm + ptol)) { S.term = TERM_PARAMETER; S.done = 1; return; }  // candidate not applied
    return;
  }
}
// after the evaluation at the candidate (cost, H, g all available in S.tot)
__device__ void lm_decide(LMState& S) {
  const double ftol = 0.1 * kSophusEps;
  const double cand_cost = S.tot[27];
  const double cost_change = S.cost - cand_cost;
  if (fabs(cost_change) <= ftol * S.cost) { S.term = TERM_FUNCTION; S.done = 1; return; }  // candidate not applied
  const double rel = cost_change / S.model;
  if (rel > 1e-3) {
    for (int i = 0; i < 7; i++) S.x[i] = S.cand[i];
    S.x_norm = norm7(S.x);
    for (int i = 0; i < 21; i++) S.H[i] = S.tot[i];
    for (int i = 0; i < 6; i++) S.g[i] = S.tot[21 + i];
    S.cost = cand_cost;
    S.gmax = grad_max_norm(S.x, S.g);
    const double t = 2.0 * rel - 1.0;
    S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
    S.radius = fmin(1e16, S.radius);
    S.decrease_factor = 2.0; S.reuse_diag = 0; S.last_successful = 1;
  } else {
    S.radius = S.radius / S.decrease_factor; S.decrease_factor *= 2.0; S.reuse_diag = 1; S.last_successful = 0;
  }
}

struct LMArgs {
  CloudView sv, tv;
  LMConfig cfg;
  const int* corr;
  const double* w;
  RegCtl* ctl;
  double* partials;       // [2][gridDim.x * kAcc]
  const double* eval_pose;  // != null: evaluate once at this pose, write kAcc totals to eval_out, return
  double* eval_out;
};

// One full sweep over the correspondence list at pose x7: every block ends with the grid totals in S.tot.
__device__ void sweep(const LMArgs& a, const double* x7, int buf, LMState& S, double (*s_red)[kAcc], cg::grid_group& grid) {
  RT P;
  quat_to_R(x7, P.R);
  P.t[0] = x7[4]; P.t[1] = x7[5]; P.t[2] = x7[6];
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; i++) acc[i] = 0.0;
  const int ncorr = a.sv.nslots * a.cfg.kc;
  const double kappa = 1.0 - a.cfg.eps;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < ncorr; r += gridDim.x * blockDim.x) {
    const int ts = __ldg(&a.corr[r]);
    if (ts < 0) continue;
    const double w = __ldg(&a.w[r]);
    if (w == 0.0) continue;
    accumulate_residual<true>(a.sv, a.tv, r / a.cfg.kc, ts, w, P, kappa, a.cfg.algo, acc);
  }
  // warp butterfly (fixed order => deterministic), then fixed-order block and grid sums
#pragma unroll
  for (int i = 0; i < kAcc; i++)
    for (int o = 16; o; o >>= 1) acc[i] += __shfl_xor_sync(kFullMask, acc[i], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < kAcc; i++) s_red[warp][i] = acc[i];
  __syncthreads();
  double* part = a.partials + (size_t)buf * gridDim.x * kAcc;
  if (threadIdx.x < kAcc) {
    double s = 0;
    for (int wv = 0; wv < kLmThreads / 32; wv++) s += s_red[wv][threadIdx.x];
    part[(size_t)blockIdx.x * kAcc + threadIdx.x] = s;
  }
  grid.sync();
  if (threadIdx.x < kAcc * 8) {
    const int comp = threadIdx.x >> 3, sub = threadIdx.x & 7;
    double s = 0;
    for (int b = sub; b < (int)gridDim.x; b += 8) s += __ldcg(&part[(size_t)b * kAcc + comp]);
    s += __shfl_xor_sync(kFullMask, s, 4);
    s += __shfl_xor_sync(kFullMask, s, 2);
    s += __shfl_xor_sync(kFullMask, s, 1);
    if (sub == 0) S.tot[comp] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kLmThreads) lm_kernel(LMArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ LMState S;
  __shared__ double s_red[kLmThreads / 32][kAcc];
  if (a.eval_pose) {  // parity-test entry: one evaluation
    if (threadIdx.x < 7) S.x[threadIdx.x] = a.eval_pose[threadIdx.x];
    __syncthreads();
    sweep(a, S.x, 0, S, s_red, grid);
    if (blockIdx.x == 0 && threadIdx.x < kAcc) a.eval_out[threadIdx.x] = S.tot[threadIdx.x];
    return;
  }
  if (threadIdx.x == 0) S.done = a.ctl->converged;
  if (threadIdx.x < 7) S.x[threadIdx.x] = a.ctl->pose[threadIdx.x];
  __syncthreads();
  if (S.done) return;
  int buf = 0, evals = 1;
  sweep(a, S.x, buf, S, s_red, grid);
  buf ^= 1;
  if (threadIdx.x == 0) { lm_init(S); lm_propose(S, a.cfg.max_iter); }
  __syncthreads();
  while (!S.done) {
    sweep(a, S.cand, buf, S, s_red, grid);
    buf ^= 1;
    evals++;
    if (threadIdx.x == 0) { lm_decide(S); if (!S.done) lm_propose(S, a.cfg.max_iter); }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // outer-loop bookkeeping: mse = |log(cur^-1 est)|^2 (impl/gicp.hpp:153), stop rule, pass trace
    RegCtl* c = a.ctl;
    double lg[6];
    pose_log(pose_mul(pose_inv(pose_from7(c->pose)), pose_from7(S.x)), lg);
    double mse = 0;
    for (int i = 0; i < 6; i++) mse += lg[i] * lg[i];
    const int before = c->outer;
    bool conv;
    if (a.cfg.algo == SICP_ALGO_SEMANTIC) conv = (mse < a.cfg.mse_stop) || (before + 1 > a.cfg.outer_cap);  // count++ first (semantic_icp.hpp:47)
    else conv = (mse < a.cfg.mse_stop) || (before > a.cfg.outer_cap);
    if (before < 64) { for (int i = 0; i < 7; i++) c->pass_pose[before][i] = S.x[i]; c->pass_lm_iters[before] = S.iter; }
    for (int i = 0; i < 7; i++) c->pose[i] = S.x[i];
    c->outer = before + 1;
    c->lm_iters_total += S.iter;
    c->lm_evals_total += evals;
    c->term_last = S.term;
    c->n_corr_last = c->n_corr_pass;
    c->n_corr_pass = 0;
    c->final_cost = S.cost;
    c->last_mse = mse;
    if (conv && !(mse < a.cfg.mse_stop)) c->flags |= 1;
    c->converged = conv ? 1 : 0;
  }
}

// ------------------------------------------------------------------ fused labels (impl/em_icp.hpp:202-268)
__global__ void fused_labels_kernel(CloudView sv, CloudView tv, double eps, double gate_d2, const double* __restrict__ pose7,
                                    const int* __restrict__ corr, const float* __restrict__ d2, uint32_t* __restrict__ labels_out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= sv.nslots) return;
  const int o = __float_as_int(sv.pts[slot].w);
  if (o < 0) return;
  const int N = sv.N;
  RT P;
  quat_to_R(pose7, P.R);
  P.t[0] = pose7[4]; P.t[1] = pose7[5]; P.t[2] = pose7[6];
  const double kappa = 1.0 - eps;
  double ps[3], ns[3];
  load_point(sv, slot, ps, ns);
  int tsl[4]; double gate[4];
  for (int c = 0; c < 4; c++) {
    const int ts = corr[(size_t)slot * 4 + c];
    tsl[c] = -1; gate[c] = 0.0;
    if (ts >= 0 && (double)d2[(size_t)slot * 4 + c] < gate_d2) {
      double pt[3], nt[3], m[3], d[3], b[3];
      load_point(tv, ts, pt, nt);
      for (int i = 0; i < 3; i++) {
        m[i] = P.R[3 * i] * ns[0] + P.R[3 * i + 1] * ns[1] + P.R[3 * i + 2] * ns[2];
        d[i] = pt[i] - (P.R[3 * i] * ps[0] + P.R[3 * i + 1] * ps[1] + P.R[3 * i + 2] * ps[2] + P.t[i]);
      }
      apply_Minv(nt, m, d, kappa, b);
      const double mahal = -0.5 * (d[0] * b[0] + d[1] * b[1] + d[2] * b[2]);
      const double two_pi = 6.283185307179586;
      const double density = pow((two_pi * two_pi * two_pi) * det_S(nt, m, kappa), -0.5) * exp(mahal);
      tsl[c] = ts;
      gate[c] = (density == 0.0) ? 0.0 : 1.0;
    }
  }
  double best = 0.0; int best_s = 0;
  const double* as = sv.avec + (size_t)slot * N;
  for (int s = 0; s < N; s++) {
    double sp = 0.0;
    for (int c = 0; c < 4; c++)
      if (tsl[c] >= 0) sp += (tv.avec[(size_t)tsl[c] * N + s] * as[s]) * gate[c];
    if (sp > best) { best = sp; best_s = s; }
  }
  labels_out[o] = (uint32_t)(best_s + 1);
}

// ------------------------------------------------------------------ host launchers
int lm_grid_blocks(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int sms = 148, per_sm = 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lm_kernel, kLmThreads, 0);
  if (per_sm < 1) per_sm = 1;
  int g = sms * std::min(per_sm, 2);
  if (device >= 0 && device < 64) cached[device] = g;
  return g;
}

sicp_status launch_estep(const sicp_cloud* src, const sicp_cloud* tgt, const LMConfig& cfg, double gate_d2, const double* d_pose7,
                         const int* d_stop, int* d_corr, const float* d_d2, double* d_w, RegCtl* d_ctl, cudaStream_t st) {
  const int n = src->nslots * cfg.kc;
  if (n == 0) return SICP_OK;
  estep_kernel<<<(n + 255) / 256, 256, 0, st>>>(src->view(), tgt->view(), cfg.algo, cfg.kc, cfg.eps, gate_d2, d_pose7, d_stop, d_corr, d_d2, d_w, d_ctl);
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}

static sicp_status launch_lm_args(LMArgs& args, int grid, cudaStream_t st) {
  void* params[] = {&args};
  SICP_CUDA(cudaLaunchCooperativeKernel((void*)lm_kernel, dim3(grid), dim3(kLmThreads), params, 0, st));
  count_launches(1);
  return SICP_OK;
}

sicp_status launch_lm(const sicp_cloud* src, const sicp_cloud* tgt, const LMConfig& cfg, const int* d_corr, const double* d_w, RegCtl* d_ctl,
                      double* d_partials, int grid, cudaStream_t st) {
  LMArgs args{src->view(), tgt->view(), cfg, d_corr, d_w, d_ctl, d_partials, nullptr, nullptr};
  return launch_lm_args(args, grid, st);
}

sicp_status launch_evaluate(const sicp_cloud* src, const sicp_cloud* tgt, const LMConfig& cfg, const int* d_corr, const double* d_w,
                            const double* d_pose7, double* d_out28, double* d_partials, int grid, cudaStream_t st) {
  LMArgs args{src->view(), tgt->view(), cfg, d_corr, d_w, nullptr, d_partials, d_pose7, d_out28};
  return launch_lm_args(args, grid, st);
}

sicp_status launch_fused_labels(const sicp_cloud* src, const sicp_cloud* tgt, double eps, double gate_d2, const double* d_pose7, const int* d_corr,
                                const float* d_d2, uint32_t* d_labels_out, cudaStream_t st) {
  if (src->nslots == 0) return SICP_OK;
  fused_labels_kernel<<<(src->nslots + 127) / 128, 128, 0, st>>>(src->view(), tgt->view(), eps, gate_d2, d_pose7, d_corr, d_d2, d_labels_out);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}

}  // namespace sicp
