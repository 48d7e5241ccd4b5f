// lm.cu — E-step weights (K3) and the device-side M-step (K4 + K5): analytic SE(3) residual/Jacobian evaluation,
// fixed-order reduction of J^T J (21) + J^T r (6) + cost (1), and a Ceres-mirroring Levenberg-Marquardt loop that
// runs entirely inside ONE cooperative kernel per outer pass (no host round trips inside an inner solve): every CTA
// sweeps its share of the residuals, a controller CTA sums the partials, takes the LM decision and publishes the next pose.
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
//   J_upsilon = -2 c,   J_omega = 2 c x (p_s + C_s c),   c = R^T M d           (verified vs the 1x7 route);
// the sweep evaluates it in the target frame and the rotation is applied once to the reduced totals (see sweep_acc).
#include <algorithm>
#include <cfloat>
#include "common.cuh"
#include "kernels.h"
#include "se3.cuh"

namespace sicp {

constexpr int kAcc = 28;  // 21 lower-triangular H + 6 g + cost
constexpr unsigned kFullMask = 0xffffffffu;

struct RT { double R[9]; double t[3]; };

__device__ __forceinline__ void load_point(const CloudView& v, int slot, double* p, double* n) {
  const float4 a = __ldg(&v.pts[slot]);
  p[0] = a.x; p[1] = a.y; p[2] = a.z;
  const double2 a01 = __ldg(reinterpret_cast<const double2*>(v.nrm) + 2 * (size_t)slot);
  n[0] = a01.x; n[1] = a01.y;
  n[2] = __ldg(&v.nrm[4 * (size_t)slot + 2]);
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

// GICPCostFunction::Probability converted to bool (gicp_cost_function.h:75-87): is
//   density = pow(det(2 pi S), -1/2) * exp(mahal)
// exactly zero in the reference's arithmetic?  exp() is rounded to a double first — into the DENORMAL range when mahal is
// in (-745.13, -708.4) — and the product is rounded again, so the answer near the bottom of the range depends on both
// roundings.  Device exp() is not trusted there: the first rounding is redone in integer units of the smallest denormal
// (2^-1074; rint = round-half-even like the hardware), the second is the comparison with half a unit.  Callers only get
// here for mahal <= -700; above that the density is a normal number times >= 0.022 and cannot vanish.
__device__ __forceinline__ bool density_is_zero(double det2pi, double mahal) {
  const double p = pow(det2pi, -0.5);
  const double units = rint(exp(mahal + 744.4400719213812623));  // exp(mahal) / 2^-1074, rounded as the reference's exp rounds it
  return p * units <= 0.5;                                        // false for NaN: a NaN density converts to `true`
}

// ------------------------------------------------------------------ K3: E-step
// Besides the weights it GATHERS the target point and normal of every candidate pair, and copies the source point and
// normal, into the group blocks of kernels.h (struct Rec), so that the many LM sweeps of the pass stream them with one
// bulk copy per 32 slots instead of repeating the gather.
// One thread per SOURCE SLOT and its KC candidates: the source row a_s, the source point/normal and the rotated
// normal are fetched / computed once per slot, and the KC independent gathers of the target rows a_t are in flight
// together (16-byte loads when the row is 16-byte aligned), which is what hides the gather latency — this kernel
// moves 261 B per candidate and computes almost nothing.
constexpr int kEstepThreads = 128;
#ifndef SICP_ESTEP_MINB
#define SICP_ESTEP_MINB 4   // resident blocks per SM the register allocation is held to (4: 128 registers, no spills; measured: 5 -> 96
                            // registers, 0.148 ms per 6 launches, 6 -> 80 registers, 0.142 ms, against 0.130 ms at 4)
#endif
template <int KC>
__global__ void __launch_bounds__(kEstepThreads, SICP_ESTEP_MINB) estep_kernel(CloudView sv, CloudView tv, int algo, double eps, double gate_d2,
                                                              const double* __restrict__ pose7, const int* __restrict__ stop,
                                                              int* __restrict__ corr, const float* __restrict__ d2,
                                                              char* __restrict__ rec, RegCtl* ctl) {
  if (stop && *stop) return;
  const long long gs = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gs < sv.nslots;
  const int slot = live ? (int)gs : 0;
  int ts[KC];
  double w[KC];
#pragma unroll
  for (int c = 0; c < KC; c++) {
    const size_t r = (size_t)slot * KC + c;
    ts[c] = live ? corr[r] : -1;
    w[c] = 0.0;
    if (ts[c] >= 0 && !((double)d2[r] < gate_d2)) { ts[c] = -1; corr[r] = -1; }  // `distSq < 250` (gicp.hpp:70, em_icp.hpp:65)
  }
  // gather the target point / normal of every kept candidate (independent loads, issued together)
  float4 tp[KC];
  double nt[KC][3];
#pragma unroll
  for (int c = 0; c < KC; c++) {
    const int t = ts[c] >= 0 ? ts[c] : 0;
    tp[c] = __ldg(&tv.pts[t]);
    const double2 n01 = __ldg(reinterpret_cast<const double2*>(tv.nrm) + 2 * (size_t)t);
    nt[c][0] = n01.x; nt[c][1] = n01.y; nt[c][2] = __ldg(&tv.nrm[4 * (size_t)t + 2]);
    if (ts[c] >= 0) w[c] = 1.0;
  }
  double ps[3], ns[3];
  load_point(sv, slot, ps, ns);
  if (algo == SICP_ALGO_EM) {
    // label compatibility (em_icp.hpp:84-89) with a_p = CM^T dist_p precomputed per point: w = a_t . a_s, summed in
    // ascending class order
    const int N = sv.N;
    const double* as = sv.avec + (size_t)slot * N;
    const double* at[KC];
    double prob[KC];
#pragma unroll
    for (int c = 0; c < KC; c++) { at[c] = tv.avec + (size_t)(ts[c] >= 0 ? ts[c] : 0) * N; prob[c] = 0.0; }
    if ((N & 1) == 0) {  // rows start on 16-byte boundaries
#pragma unroll 2
      for (int s = 0; s < N; s += 2) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(as + s));
#pragma unroll
        for (int c = 0; c < KC; c++) {
          const double2 t = __ldg(reinterpret_cast<const double2*>(at[c] + s));
          prob[c] += t.x * a.x;
          prob[c] += t.y * a.y;
        }
      }
    } else {
      for (int s = 0; s < N; s++) {
        const double a = __ldg(&as[s]);
#pragma unroll
        for (int c = 0; c < KC; c++) prob[c] += __ldg(&at[c][s]) * a;
      }
    }
    // Probability() -> bool (gicp_cost_function.h:75-87): weight kept iff the density is not exactly 0
    RT P;
    quat_to_R(pose7, P.R);
    P.t[0] = pose7[4]; P.t[1] = pose7[5]; P.t[2] = pose7[6];
    double m[3], q[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      m[i] = P.R[3 * i] * ns[0] + P.R[3 * i + 1] * ns[1] + P.R[3 * i + 2] * ns[2];
      q[i] = P.R[3 * i] * ps[0] + P.R[3 * i + 1] * ps[1] + P.R[3 * i + 2] * ps[2] + P.t[i];
    }
    const double kappa = 1.0 - eps;
#pragma unroll
    for (int c = 0; c < KC; c++) {
      if (ts[c] < 0) continue;
      const double d[3] = {(double)tp[c].x - q[0], (double)tp[c].y - q[1], (double)tp[c].z - q[2]};
      double b[3];
      apply_Minv(nt[c], m, d, kappa, b);
      const double mahal = -0.5 * (d[0] * b[0] + d[1] * b[1] + d[2] * b[2]);
      // The density can only underflow to exactly 0 when exp(mahal) is near the bottom of the double range:
      // det(2 pi S) <= (2 pi)^3 * 8, so the pow factor is >= 0.022 and for mahal > -700 the product is >= 1e-306.
      // Only candidates beyond that (Mahalanobis^2 > 1400) pay for the pow / exp of the reference expression.
      if (!(mahal > -700.0)) {
        const double two_pi = 6.283185307179586;
        const double det2pi = (two_pi * two_pi * two_pi) * det_S(nt[c], m, kappa);
        if (density_is_zero(det2pi, mahal)) prob[c] *= 0.0;  // NaN stays "true" like the bool conversion
      }
      w[c] = prob[c];
    }
  }
  int kept = 0;
  if (live) {
    char* blk = rec + (size_t)(slot >> 5) * Rec::group_bytes(KC);
    const int lane = slot & 31;
#pragma unroll
    for (int c = 0; c < KC; c++) {
      char* cb = blk + c * Rec::kCandBytes;
      const bool ok = ts[c] >= 0;
      kept += ok;
      reinterpret_cast<double*>(cb + Rec::kW)[lane] = w[c];
      // no residual: zeroed geometry keeps the branch-free sweep finite (its weight is 0)
      reinterpret_cast<float*>(cb + Rec::kPx)[lane] = ok ? tp[c].x : 0.f;
      reinterpret_cast<float*>(cb + Rec::kPy)[lane] = ok ? tp[c].y : 0.f;
      reinterpret_cast<float*>(cb + Rec::kPz)[lane] = ok ? tp[c].z : 0.f;
      reinterpret_cast<double*>(cb + Rec::kNx)[lane] = ok ? nt[c][0] : 0.0;
      reinterpret_cast<double*>(cb + Rec::kNy)[lane] = ok ? nt[c][1] : 0.0;
      reinterpret_cast<double*>(cb + Rec::kNz)[lane] = ok ? nt[c][2] : 0.0;
    }
    // source side; padding slots (NaN coordinates, undefined normals) and slots without any residual are zeroed
    char* sb = blk + KC * Rec::kCandBytes;
    const bool any = kept > 0;
    reinterpret_cast<float*>(sb + Rec::kSx)[lane] = any ? (float)ps[0] : 0.f;
    reinterpret_cast<float*>(sb + Rec::kSy)[lane] = any ? (float)ps[1] : 0.f;
    reinterpret_cast<float*>(sb + Rec::kSz)[lane] = any ? (float)ps[2] : 0.f;
    reinterpret_cast<double*>(sb + Rec::kSnx)[lane] = any ? ns[0] : 0.0;
    reinterpret_cast<double*>(sb + Rec::kSny)[lane] = any ? ns[1] : 0.0;
    reinterpret_cast<double*>(sb + Rec::kSnz)[lane] = any ? ns[2] : 0.0;
  }
  if (ctl) {  // residual blocks of this pass (diagnostics)
    kept += __shfl_xor_sync(kFullMask, kept, 16); kept += __shfl_xor_sync(kFullMask, kept, 8); kept += __shfl_xor_sync(kFullMask, kept, 4);
    kept += __shfl_xor_sync(kFullMask, kept, 2); kept += __shfl_xor_sync(kFullMask, kept, 1);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&ctl->n_corr_pass, kept);
  }
}

// ------------------------------------------------------------------ K4 + K5: LM
// Solver state of one inner solve.  It lives in the shared memory of the controller block (block 0) for the whole
// launch; LMSync in global memory carries only the arrival counter, the generation flag and the broadcast pose.
struct LMState {
  double x[7], cand[7];
  double Hs[36], gs[6];        // Jacobi-scaled Gauss-Newton system at x
  double g[6], cost;
  double scale[6], diag[6];
  double radius, decrease_factor, x_norm, gmax, model;
  int reuse_diag, last_successful, invalid, iter, evals, term, done, started;
  int pad0, pad1;
};
struct LMSync {
  unsigned count;              // blocks that finished the current sweep
  unsigned flag;               // generation of the last published control step
  unsigned pad[2];
  double bcast[8];             // candidate pose [7] + done
};
enum { TERM_NO_CONV = 0, TERM_GRADIENT = 1, TERM_PARAMETER = 2, TERM_FUNCTION = 3, TERM_RADIUS = 4, TERM_FAIL = 5 };

__device__ __forceinline__ int tri(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// ---- branch-free FP64 special functions for the sweep.  The library versions carry slow-path subroutine calls for
// denormal / infinite / negative arguments, which split the residual code into many basic blocks and stop the
// compiler from interleaving independent residuals; here the arguments are known to be normal and positive
// (sum >= 1, v >= 2.2e-16, det in (0, 4]), so a MUFU seed + two Newton steps (full double accuracy) is enough.
__device__ __forceinline__ double rcp_pos(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  e = fma(-a, x, 1.0);
  return fma(x, e, x);
}
__device__ __forceinline__ double rsqrt_pos(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  return fma(y, e, y);
}
// One third-order step instead of two Newton steps.  The MUFU seeds read only the top 20 mantissa bits of the argument, so
// their relative error e is <= ~2^-20 and the remainder e^3 ~ 2^-60 is below the rounding of the result: same accuracy
// class (<= 1 ulp) for one FP64 instruction less (rcp) or two less (rsqrt).  Used by the residual sweep only.
__device__ __forceinline__ double rcp_pos3(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double e = fma(-a, x, 1.0);          // 1/a = x / (1 - e) = x (1 + e + e^2 + ...)
  return fma(x, fma(e, e, e), x);
}
__device__ __forceinline__ double rsqrt_pos3(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-(a * y), y, 1.0);    // a y^2 = 1 - e;  a^-1/2 = y (1 - e)^-1/2 = y (1 + e/2 + 3 e^2/8 + ...)
  return fma(y, e * fma(0.375, e, 0.5), y);
}
// log(x) for normal x >= 1 (fdlibm e_log.c reduction and minimax coefficients; error < 1 ulp)
__device__ __forceinline__ double log_ge1(double x) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int k = (hi >> 20) - 1023;
  hi &= 0x000fffff;
  const int i = (hi + 0x95f64) & 0x100000;       // mantissa above sqrt(2): halve it, bump the exponent
  k += i >> 20;
  const double m = __hiloint2double(hi | (i ^ 0x3ff00000), lo);
  const double f = m - 1.0;
  const double sden = rcp_pos(2.0 + f);
  const double ss = f * sden;
  const double z = ss * ss;
  const double w = z * z;
  const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
  const double t2 = z * (6.666666666666735130e-01 + w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  // log(1+f) = f - (hfsq - s*(hfsq+R));  log(x) = k*ln2_hi + (log(1+f) + k*ln2_lo)
  return dk * 6.93147180369123816490e-01 - ((hfsq - (ss * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
}

// log(x) for normal x >= 1 through a 256-entry table in shared memory (fill_log_table): x = 2^k m, m in [1, 2); entry i
// covers m in [1 + i/256, 1 + (i+1)/256) with c_i = 1 / midpoint and L_i = -log(c_i), so log(m) = L_i + log1p(r),
// r = m c_i - 1, |r| <= 2^-9, and a degree-6 series is exact to 2^-63 / 7.  12 FP64 instructions against the 26 of
// log_ge1 (no reciprocal, short polynomial): the loss is a third of the sweep's FP64 work.  Absolute error ~1e-16.
#ifndef SICP_LM_LOGTAB
#define SICP_LM_LOGTAB 1
#endif
constexpr int kLogTab = 256;
__device__ __forceinline__ void fill_log_table(double2* tab, int tid, int nthreads) {
  for (int i = tid; i < kLogTab; i += nthreads) {
    const double c = 1.0 / (1.0 + (i + 0.5) * (1.0 / kLogTab));
    tab[i] = make_double2(c, -log(c));
  }
}
__device__ __forceinline__ double log_tab(double x, const double2* __restrict__ tab) {
  const int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  const int k = (hi >> 20) - 1023;
  const double2 t = tab[(hi >> 12) & (kLogTab - 1)];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double r = fma(m, t.x, -1.0);
  double q = fma(r, -1.0 / 6.0, 0.2);
  q = fma(r, q, -0.25);
  q = fma(r, q, 1.0 / 3.0);
  q = fma(r, q, -0.5);
  const double dk = (double)k;
  const double small = fma(r * r, q, fma(dk, 1.90821492927058770002e-10, r));  // r^2 q + k ln2_lo + r
  return fma(dk, 6.93147180369123816490e-01, t.y + small);                    // k ln2_hi is exact (21 trailing zero bits)
}

// T * exp(delta) (local_parameterization_se3.h:22) on the LM critical path: one sincos, reciprocal multiplies and an
// rsqrt renormalisation instead of the divisions / square roots of the general-purpose se3.cuh routines.
__device__ __forceinline__ void pose_plus_fast(const double* x7, const double* d, double* out7) {
  const double* up = d;
  const double* om = d + 3;
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double imag, real, a, b;  // exp: q = (imag*om, real), V = I + a*Om + b*Om^2
  if (th2 < 1e-16) {
    imag = 0.5 - (1.0 / 48.0) * th2;
    real = 1.0 - (1.0 / 8.0) * th2;
    a = 0.5 - th2 * (1.0 / 24.0);
    b = (1.0 / 6.0) - th2 * (1.0 / 120.0);
  } else {
    const double inv_th = rsqrt(th2);
    const double theta = th2 * inv_th;
    double sh, ch;
    sincos(0.5 * theta, &sh, &ch);
    imag = sh * inv_th;
    real = ch;
    const double s_th = 2.0 * sh * ch;          // sin(theta)
    const double omc = 2.0 * sh * sh;           // 1 - cos(theta)
    const double inv_th2 = inv_th * inv_th;
    a = omc * inv_th2;
    b = (theta - s_th) * inv_th2 * inv_th;
  }
  const double qe[4] = {imag * om[0], imag * om[1], imag * om[2], real};
  // t_e = V * upsilon, with Om*u = om x u and Om^2*u = om x (om x u)
  double c1[3], c2[3], te[3];
  cross3(om, up, c1);
  cross3(om, c1, c2);
#pragma unroll
  for (int i = 0; i < 3; i++) te[i] = up[i] + a * c1[i] + b * c2[i];
  // compose: q = q_x * q_e (renormalised), t = t_x + q_x * t_e
  const double ax = x7[0], ay = x7[1], az = x7[2], aw = x7[3];
  double q[4];
  q[3] = aw * qe[3] - ax * qe[0] - ay * qe[1] - az * qe[2];
  q[0] = aw * qe[0] + ax * qe[3] + ay * qe[2] - az * qe[1];
  q[1] = aw * qe[1] + ay * qe[3] + az * qe[0] - ax * qe[2];
  q[2] = aw * qe[2] + az * qe[3] + ax * qe[1] - ay * qe[0];
  const double inv_n = rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
#pragma unroll
  for (int i = 0; i < 4; i++) out7[i] = q[i] * inv_n;
  double rt[3];
  quat_rot(x7, te, rt);
#pragma unroll
  for (int i = 0; i < 3; i++) out7[4 + i] = x7[4 + i] + rt[i];
}

// Ceres' gradient_max_norm = |x - Plus(x, -g)|_inf, compared with 1e-11.  The exp map is only evaluated when the
// gradient is small enough for that test to possibly pass (|delta| >= 1e-6 moves some coordinate by >> 1e-11).
__device__ __forceinline__ double grad_max_norm(const double* x7, const double* g) {
  double gm = 0;
#pragma unroll
  for (int j = 0; j < 6; j++) gm = fmax(gm, fabs(g[j]));
  if (gm > 1e-6) return INFINITY;
  double ng[6];
  for (int j = 0; j < 6; j++) ng[j] = -g[j];
  double p7[7];
  pose_plus_fast(x7, ng, p7);
  double m = 0;
  for (int i = 0; i < 7; i++) m = fmax(m, fabs(x7[i] - p7[i]));
  return m;
}
__device__ __forceinline__ double norm7(const double* a) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) s += a[i] * a[i];
  return sqrt(s);
}
// Warp-cooperative form of lm_control (same decisions, same state), run by warp 0 of the controller block.
// The single-thread version is bound by latency: ~500 dependent FP64 ops at 8 cycles plus a 29-cycle shared-memory
// round trip for every access to the solver state.  Here every lane loads the scalar state into registers ONCE, all
// lanes make the (uniform) decisions redundantly, and the 6x6 work is spread over lanes: lane i owns row i of the
// scaled Gauss-Newton system (formed in registers), of the Cholesky factor and of the triangular solves (pivots and
// solution components travel by shuffle); the model decrease is a 6-lane dot product.  State goes back to shared
// memory once, at the end.
__device__ __forceinline__ void lm_control_warp(LMState& S, const double* tot, int max_iter, double* s_L, int lane,
                                                const double ftol = 0.1 * kSophusEps, const double gtol = 0.1 * kSophusEps) {
  const double ptol = 1e-8;
  const int i = lane < 6 ? lane : 0;  // lanes >= 6 shadow row 0; nothing is ever read from them
  // ---- state -> registers
  double x[7], cand[7], scale[6], Hrow[6];
#pragma unroll
  for (int q = 0; q < 7; q++) { x[q] = S.x[q]; cand[q] = S.cand[q]; }
#pragma unroll
  for (int j = 0; j < 6; j++) { scale[j] = S.scale[j]; Hrow[j] = S.Hs[6 * i + j]; }
  double gs_i = S.gs[i], diag_i = S.diag[i];
  double radius = S.radius, decrease = S.decrease_factor, x_norm = S.x_norm, gmax = S.gmax, model = S.model, cost = S.cost;
  int reuse_diag = S.reuse_diag, last_successful = S.last_successful, invalid = S.invalid, iter = S.iter, term = S.term, done = 0;
  const int started = S.started;
  __syncwarp();  // every lane holds its copy before any lane stores back (the stores are at the end and in the adopt step)
  // ---- iteration zero, or accept / reject the evaluated candidate (uniform)
  int adopt = 0;
  if (!started) {
    radius = 1e4; decrease = 2.0; reuse_diag = 0; last_successful = 1; invalid = 0; iter = 0; term = TERM_NO_CONV;
    adopt = 1;
  } else {
    const double cost_change = cost - tot[27];
    if (fabs(cost_change) <= ftol * cost) { term = TERM_FUNCTION; done = 1; }  // candidate not applied
    else {
      const double rel = cost_change / model;
      if (rel > 1e-3) {
#pragma unroll
        for (int q = 0; q < 7; q++) x[q] = cand[q];
        const double t = 2.0 * rel - 1.0;
        radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
        decrease = 2.0; reuse_diag = 0; last_successful = 1;
        adopt = 2;
      } else {
        radius = radius / decrease; decrease *= 2.0; reuse_diag = 1; last_successful = 0;
      }
    }
  }
  // ---- adopt the evaluation in `tot` as the linearisation point (lm_adopt)
  if (adopt && !done) {
    double g[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      g[j] = tot[21 + j];
      if (adopt == 1) scale[j] = 1.0 / (1.0 + sqrt(tot[tri(j, j)]));  // Jacobi scaling, computed once
    }
    double scale_i = scale[0], g_i = g[0];
#pragma unroll
    for (int j = 1; j < 6; j++) { scale_i = (i == j) ? scale[j] : scale_i; g_i = (i == j) ? g[j] : g_i; }
#pragma unroll
    for (int j = 0; j < 6; j++) Hrow[j] = tot[tri(i, j)] * scale_i * scale[j];
    gs_i = g_i * scale_i;
    cost = tot[27];
    gmax = grad_max_norm(x, g);
    x_norm = norm7(x);
    if (lane < 6) S.g[lane] = g_i;
  }
  // ---- propose the next candidate (lm_propose); repeats only after an invalid step
  while (!done) {
    if (iter >= max_iter) { term = TERM_NO_CONV; done = 1; break; }
    if (last_successful && gmax <= gtol) { term = TERM_GRADIENT; done = 1; break; }
    if (radius <= 1e-32) { term = TERM_RADIUS; done = 1; break; }
    iter++;
    if (!reuse_diag) {
      double h = Hrow[0];
#pragma unroll
      for (int j = 1; j < 6; j++) h = (i == j) ? Hrow[j] : h;
      diag_i = fmin(fmax(h, 1e-6), 1e32);
    }
    reuse_diag = 1;
    // Cholesky of (Hs + diag/radius), row i in this lane
    const double dmp = diag_i / radius;
    double A[6], Lr[6], inv[6];
#pragma unroll
    for (int j = 0; j < 6; j++) A[j] = Hrow[j] + (i == j ? dmp : 0.0);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const double piv = __shfl_sync(kFullMask, A[j], j);
      ok = ok && (piv > 0);
      inv[j] = rsqrt_pos(piv);  // pivots of an accepted factorisation are positive normal numbers; otherwise `ok` is false
      Lr[j] = A[j] * inv[j];
#pragma unroll
      for (int k = j + 1; k < 6; k++) A[k] -= Lr[j] * __shfl_sync(kFullMask, Lr[j], k);
    }
    // forward substitution L y = gs (every lane ends up with all of y)
    double r = gs_i, y[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      y[j] = __shfl_sync(kFullMask, r, j) * inv[j];
      r -= Lr[j] * y[j];
    }
    // back substitution L^T s = y needs columns of L: through shared memory
    __syncwarp();
    if (lane < 6)
#pragma unroll
      for (int j = 0; j < 6; j++) s_L[6 * lane + j] = Lr[j];
    __syncwarp();
    double sv = y[0];
#pragma unroll
    for (int j = 1; j < 6; j++) sv = (i == j) ? y[j] : sv;
    double sol[6];
#pragma unroll
    for (int k = 5; k >= 0; k--) {
      sol[k] = __shfl_sync(kFullMask, sv, k) * inv[k];
      sv -= s_L[6 * k + i] * sol[k];
    }
    // model decrease  sol.gs - sol.Hs.sol / 2  (step = -sol)
    double row = 0, si = sol[0];
#pragma unroll
    for (int j = 0; j < 6; j++) { row += Hrow[j] * sol[j]; si = (i == j) ? sol[j] : si; }
    double tsum = lane < 6 ? si * (gs_i - 0.5 * row) : 0.0;
    tsum += __shfl_xor_sync(kFullMask, tsum, 1);
    tsum += __shfl_xor_sync(kFullMask, tsum, 2);
    tsum += __shfl_xor_sync(kFullMask, tsum, 4);
    const double mdl = __shfl_sync(kFullMask, tsum, 0);
    if (!ok || !(mdl > 0)) {
      if (++invalid >= 5) { term = TERM_FAIL; done = 1; break; }
      radius /= decrease; decrease *= 2.0; last_successful = 0;
      continue;
    }
    invalid = 0;
    model = mdl;
    double delta[6];
#pragma unroll
    for (int j = 0; j < 6; j++) delta[j] = -sol[j] * scale[j];
    pose_plus_fast(x, delta, cand);
    double sn = 0;
#pragma unroll
    for (int q = 0; q < 7; q++) sn += (x[q] - cand[q]) * (x[q] - cand[q]);
    if (sqrt(sn) <= ptol * (x_norm + ptol)) { term = TERM_PARAMETER; done = 1; }  // candidate not applied
    break;
  }
  // ---- registers -> state
  if (lane < 6) {
#pragma unroll
    for (int j = 0; j < 6; j++) S.Hs[6 * lane + j] = Hrow[j];
    S.gs[lane] = gs_i; S.diag[lane] = diag_i;
    double sc = scale[0];
#pragma unroll
    for (int j = 1; j < 6; j++) sc = (lane == j) ? scale[j] : sc;
    S.scale[lane] = sc;
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 7; q++) { S.x[q] = x[q]; S.cand[q] = cand[q]; }
    S.radius = radius; S.decrease_factor = decrease; S.x_norm = x_norm; S.gmax = gmax; S.model = model; S.cost = cost;
    S.reuse_diag = reuse_diag; S.last_successful = last_successful; S.invalid = invalid; S.iter = iter; S.term = term; S.done = done;
    S.started = 1; S.evals++;
  }
  __syncwarp();
}

#ifdef SICP_STATS
static __device__ unsigned long long g_lm_blk[512][2];  // per block: cycles in sweep+reduce, cycles waiting for the next pose (summed over evals)
#endif
struct LMArgs {
  CloudView sv;
  LMConfig cfg;
  const char* rec;          // group blocks of the pass (kernels.h: struct Rec)
  RegCtl* ctl;
  double* partials;         // [gridDim.x * kAcc]
  LMSync* sync;
  const double* eval_pose;  // != null: evaluate once at this pose, write kAcc totals to eval_out, return
  double* eval_out;
  cudaGraphConditionalHandle cond;  // != 0: the launch is the last node of a graph WHILE body (register.cu); set to "not converged"
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// rho(s), rho'(s) at s = res^2 of the three Ceres loss compositions, multiplied by the weight w (the
// E-step weight for EM; exactly 1.0 or 0.0 for GICP / SemanticICP, where 0 marks "no residual in this record").
template <int ALGO>
__device__ __forceinline__ void loss_fast(double w, double res, const double2* __restrict__ logtab, double* rho0, double* rho1) {
#if SICP_LM_LOGTAB
#define SICP_LOG_GE1(x) log_tab((x), logtab)
#else
#define SICP_LOG_GE1(x) log_ge1(x)
#endif
  const double s = res * res;
  if (ALGO == SICP_ALGO_SEMANTIC) {  // CauchyLoss(1.5)
    const double sum = 1.0 + s * (1.0 / 2.25);
    *rho0 = w * (2.25 * SICP_LOG_GE1(sum));
    *rho1 = w * fmax(DBL_MIN, rcp_pos3(sum));
    return;
  }
  // ComposedLoss(CauchyLoss(3.0) [ScaledLoss w for EM], SQLoss): g = sqrt(s + eps), f = 9 log(1 + g/9)
  const double v = s + DBL_EPSILON;
  const double rs = rsqrt_pos3(v);
  const double g0 = v * rs;
  const double sum = 1.0 + g0 * (1.0 / 9.0);
  const double f0 = w * (9.0 * SICP_LOG_GE1(sum)), f1 = w * fmax(DBL_MIN, rcp_pos3(sum));
  *rho0 = f0;
  *rho1 = f1 * (0.5 * rs);
}

// ---- bulk (TMA) copies global -> shared with mbarrier completion: one elected lane arms the barrier with the byte
// count and issues the copy; every lane of the warp then waits on the barrier's phase parity
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of the buffer are done
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

// Per-warp double buffer of group blocks.  `uses` counts the blocks consumed so far over the whole launch: use u lives in
// buffer u & 1 and completes phase (u >> 1) & 1 of that buffer's barrier.
struct GroupPipe {
  char* buf;               // 2 * group_bytes of shared memory owned by this warp
  unsigned bar;            // shared address of its two 8-byte barriers
  unsigned uses;
  const char* primed_src;  // != null: the copy for use `uses` is already in flight from this address (issued as the previous sweep ended)
};

// Residual sweep at the pose (P.R, P.t): one thread owns one source slot and its KC records.
// Everything is evaluated in the TARGET frame so that no per-residual R^T products are needed:
//   d = p_t - (R p_s + t),  m = R n_s,  b = (2I - kappa(n_t n_t^T + m m^T))^-1 d,  res = d.b
//   local 6-dof Jacobian of res for T*exp(delta):  J = -2 D j,  j = [b ; v x b],  v = R p_s - kappa (m.b) m,
//   D = blockdiag(R^T, R^T)                     (equals J_ups = -2c, J_om = 2c x (p_s + C_s c), c = R^T b)
// The thread carries 2j (b2 = 2b: one FMA less per component, exact scaling) and accumulates  A_H += rho' (2j)(2j)^T
// (lower triangle),  A_g += rho' res (2j),  A_c += rho;  the remaining constants (sign of g, 1/2 of the cost) and the
// rotation D are applied ONCE to the 28 grid totals by the controller block (rotate_totals).
// The KC records of a slot are evaluated by straight-line, branch-free code (records without a residual have w = 0 and
// zeroed geometry) so that their dependency chains interleave.
// Data movement: a warp-iteration covers one GROUP of 32 slots, whose records are one contiguous block (struct Rec).
// The block of the NEXT iteration is fetched by a bulk copy into the other half of the warp's double buffer while the
// current one is evaluated out of shared memory, so the FP64 pipe never waits on an L2 round trip; the first block of
// the next sweep is fetched as this sweep ends (the records do not change during a solve) and is already waiting when
// the controller publishes the next pose.
// Work split of a sweep: a block owns a contiguous range [g0, g1) of the groups and deals them to its warps round-robin,
// so the four schedulers of an SM carry the same number of groups (+-1).  `worker` of `nworkers` equal shares; in the
// single-problem kernel the controller block (block 0) takes ctl_share8 / 8 of a normal share (0 = it does not sweep).
__device__ __forceinline__ void group_range(int ngroups, long long u0, long long u1, long long units, int* g0, int* g1) {
  *g0 = (int)(((long long)ngroups * u0) / units);
  *g1 = (int)(((long long)ngroups * u1) / units);
}
// address of the first group block this warp reads in a sweep over [g0, g1) of `rec` (null: none)
template <int KC>
__device__ __forceinline__ const char* first_block(const char* rec, int g0, int g1) {
  const int gfirst = g0 + (threadIdx.x >> 5);
  return gfirst < g1 ? rec + (size_t)gfirst * Rec::group_bytes(KC) : nullptr;
}
// next_src: first block of the sweep this warp will run NEXT (same problem, or the other one of a pair): prefetched as this
// sweep ends.
template <int ALGO, int KC, int THREADS>
__device__ __forceinline__ void sweep_acc(const LMArgs& a, const double* s_RT, const double2* __restrict__ logtab, double* acc, GroupPipe& pp, int g0, int g1,
                                          const char* next_src) {
#pragma unroll
  for (int i = 0; i < kAcc; i++) acc[i] = 0.0;
  constexpr unsigned GB = Rec::group_bytes(KC);
  constexpr int W = THREADS / 32;
  const int lane = threadIdx.x & 31;
  const int gfirst = g0 + (threadIdx.x >> 5);
  const char* rec = a.rec;
  const char* want = gfirst < g1 ? rec + (size_t)gfirst * GB : nullptr;
  if (pp.primed_src != want) {
    if (pp.primed_src) {  // a block of another sweep is in flight (the other problem of a pair finished meanwhile): let it land, drop it
      mbar_wait(pp.bar + 8 * (pp.uses & 1), (pp.uses >> 1) & 1);
      pp.uses++;
      __syncwarp();
    }
    if (want && lane == 0) bulk_load(smem_u32(pp.buf + (pp.uses & 1) * GB), want, GB, pp.bar + 8 * (pp.uses & 1));
  }
  pp.primed_src = nullptr;
  const double aa2 = a.cfg.aa * a.cfg.aa;
  for (int g = gfirst; g < g1; g += W) {
    __syncwarp();  // every lane is done with the buffer the next copy lands in (it was read one iteration ago)
    if (lane == 0 && g + W < g1) bulk_load(smem_u32(pp.buf + ((pp.uses + 1) & 1) * GB), rec + (size_t)(g + W) * GB, GB, pp.bar + 8 * ((pp.uses + 1) & 1));
    mbar_wait(pp.bar + 8 * (pp.uses & 1), (pp.uses >> 1) & 1);
    const char* blk = pp.buf + (pp.uses & 1) * GB;
    pp.uses++;
    const char* sb = blk + KC * Rec::kCandBytes;
    const double ps[3] = {(double)reinterpret_cast<const float*>(sb + Rec::kSx)[lane], (double)reinterpret_cast<const float*>(sb + Rec::kSy)[lane],
                          (double)reinterpret_cast<const float*>(sb + Rec::kSz)[lane]};
    const double ns[3] = {reinterpret_cast<const double*>(sb + Rec::kSnx)[lane], reinterpret_cast<const double*>(sb + Rec::kSny)[lane],
                          reinterpret_cast<const double*>(sb + Rec::kSnz)[lane]};
    double q0[3], qt[3], m[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {  // R, t are read from shared memory (broadcast) instead of pinning 24 registers
      const double r0 = s_RT[3 * i], r1 = s_RT[3 * i + 1], r2 = s_RT[3 * i + 2];
      q0[i] = r0 * ps[0] + r1 * ps[1] + r2 * ps[2];
      qt[i] = q0[i] + s_RT[9 + i];
      m[i] = r0 * ns[0] + r1 * ns[1] + r2 * ns[2];
    }
#pragma unroll
    for (int c = 0; c < KC; c++) {
      const char* cb = blk + c * Rec::kCandBytes;
      const double w = reinterpret_cast<const double*>(cb + Rec::kW)[lane];
      const double u[3] = {reinterpret_cast<const double*>(cb + Rec::kNx)[lane], reinterpret_cast<const double*>(cb + Rec::kNy)[lane],
                           reinterpret_cast<const double*>(cb + Rec::kNz)[lane]};
      const double d[3] = {(double)reinterpret_cast<const float*>(cb + Rec::kPx)[lane] - qt[0],
                           (double)reinterpret_cast<const float*>(cb + Rec::kPy)[lane] - qt[1],
                           (double)reinterpret_cast<const float*>(cb + Rec::kPz)[lane] - qt[2]};
      // b = M d (rank-2 Woodbury, see apply_Minv)
      const double cuv = u[0] * m[0] + u[1] * m[1] + u[2] * m[2];
      const double pu = u[0] * d[0] + u[1] * d[1] + u[2] * d[2];
      const double pm = m[0] * d[0] + m[1] * d[1] + m[2] * d[2];
      // carried as b2 = 2 b (scaling by 2 is exact): b2 = d + 2 g1 u + 2 g2 m is two FMAs per component, and the factor
      // moves into the constants (hk = 2 k4 here, hk = kappa / 2 for m.b below, 1 and -1 instead of 4 and -2 in rotate_totals)
      const double be = a.cfg.hk * cuv;
      const double idet = a.cfg.hk * rcp_pos3(fma(-be, be, aa2));          // (aa - be)(aa + be) = aa^2 - be^2
      const double g1c = (a.cfg.aa * pu + be * pm) * idet, g2c = (be * pu + a.cfg.aa * pm) * idet;
      double b[3];
#pragma unroll
      for (int i = 0; i < 3; i++) b[i] = fma(g1c, u[i], fma(g2c, m[i], d[i]));
      const double res = 0.5 * (d[0] * b[0] + d[1] * b[1] + d[2] * b[2]);
      double rho0, rho1;
      loss_fast<ALGO>(w, res, logtab, &rho0, &rho1);
      const double mb = a.cfg.hk * (m[0] * b[0] + m[1] * b[1] + m[2] * b[2]);
      const double v[3] = {q0[0] - mb * m[0], q0[1] - mb * m[1], q0[2] - mb * m[2]};
      double j[6], jw[6];
      j[0] = b[0]; j[1] = b[1]; j[2] = b[2];
      j[3] = v[1] * b[2] - v[2] * b[1];
      j[4] = v[2] * b[0] - v[0] * b[2];
      j[5] = v[0] * b[1] - v[1] * b[0];
#pragma unroll
      for (int p = 0; p < 6; p++) jw[p] = rho1 * j[p];
      int k = 0;
#pragma unroll
      for (int p = 0; p < 6; p++) {
#pragma unroll
        for (int q = 0; q <= p; q++) acc[k++] += jw[p] * j[q];
        acc[21 + p] += jw[p] * res;
      }
      acc[27] += rho0;
    }
  }
  __syncwarp();
  if (next_src && lane == 0) bulk_load(smem_u32(pp.buf + (pp.uses & 1) * GB), next_src, GB, pp.bar + 8 * (pp.uses & 1));
  pp.primed_src = next_src;
}
// a block may only exit (and hand its shared memory back) once no bulk copy is in flight into it
template <int KC, int THREADS>
__device__ __forceinline__ void pipe_drain(GroupPipe& pp) {
  if (pp.primed_src) mbar_wait(pp.bar + 8 * (pp.uses & 1), (pp.uses >> 1) & 1);
  pp.primed_src = nullptr;
}

// Block reduction of the 28 per-thread sums through shared memory (fixed order => deterministic): every thread
// stores its 28 values, then 28 x 8 threads each add one 32-value segment of one row (rotated start: conflict-free)
// and a 3-step butterfly joins the 8 segments.  Leaves the block's sums in part[blockIdx.x][0..27].
template <int THREADS>
__device__ __forceinline__ void block_reduce(const double* acc, double* s_acc, double* part) {
  constexpr int kLmWarps = THREADS / 32;
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < kAcc; i++) s_acc[i * THREADS + tid] = acc[i];
  __syncthreads();
  {  // kAcc * kLmWarps <= THREADS (kAcc < 32); every thread runs the shuffles, the surplus ones carry zeros
    const bool active = tid < kAcc * kLmWarps;
    const int row = active ? tid / kLmWarps : 0, seg = tid % kLmWarps;
    const double* base = s_acc + row * THREADS + seg * 32;
    double s = 0;
    if (active) {
#pragma unroll 8
      for (int k = 0; k < 32; k++) s += base[(k + tid) & 31];
    }
#pragma unroll
    for (int off = 1; off < kLmWarps; off <<= 1) s += __shfl_xor_sync(kFullMask, s, off);  // kLmWarps is a power of two
    if (active && seg == 0) part[(size_t)blockIdx.x * kAcc + row] = s;
  }
}

// Fixed-order sum of the block partials by the controller block: warp `sub` sums blocks sub, sub+8, ... with
// independent loads, then the 8 warp sums are added in order.  Result in s_tot[0..27].
template <int THREADS>
__device__ __forceinline__ void reduce_partials(const double* part, double (*s_red)[kAcc], double* s_tot, int b0 = 0) {
  constexpr int kLmWarps = THREADS / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0;
  if (lane < kAcc) {
    constexpr int kMaxPerWarp = 320 / kLmWarps;  // grids of up to 320 blocks
    double v[kMaxPerWarp];
#pragma unroll
    for (int k = 0; k < kMaxPerWarp; k++) {
      const int bI = b0 + warp + kLmWarps * k;
      v[k] = bI < (int)gridDim.x ? __ldcg(&part[(size_t)bI * kAcc + lane]) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kMaxPerWarp; k++) s += v[k];
  }
  if (lane < kAcc) s_red[warp][lane] = s;
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double t = 0;
#pragma unroll
    for (int wv = 0; wv < kLmWarps; wv++) t += s_red[wv][threadIdx.x];
    s_tot[threadIdx.x] = t;
  }
  __syncthreads();
}

// Grid totals (A_H, A_g, A_c) -> Ceres quantities in the local frame of the evaluated pose:
//   H = D A_H D^T,  g = -D A_g,  cost = A_c / 2,  D = blockdiag(R^T, R^T)  (the sweep accumulates with 2j, so the 4 and -2
//   of J = -2 D j are already inside A_H and A_g).  28 threads, one output each.
__device__ __forceinline__ void rotate_totals(const double* s_tot, const double* R, double* s_rot) {
  const int e = threadIdx.x;
  if (e < 21) {
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= e) a++;
    const int b = e - a * (a + 1) / 2;
    const int A = a / 3, i = a % 3, B = b / 3, jx = b % 3;
    double s = 0;
    for (int k = 0; k < 3; k++)
      for (int l = 0; l < 3; l++) s += R[3 * k + i] * s_tot[tri(3 * A + k, 3 * B + l)] * R[3 * l + jx];
    s_rot[e] = s;
  } else if (e < 27) {
    const int a = e - 21, A = a / 3, i = a % 3;
    double s = 0;
    for (int k = 0; k < 3; k++) s += R[3 * k + i] * s_tot[21 + 3 * A + k];
    s_rot[e] = -s;
  } else if (e == 27) {
    s_rot[27] = 0.5 * s_tot[27];
  }
}

// Outer-loop bookkeeping when an inner solve has terminated (one thread): mse = |log(cur^-1 est)|^2 (impl/gicp.hpp:153),
// the stopping rule, the pass trace.  Returns `converged`.
template <int ALGO>
__device__ __forceinline__ bool finish_pass(RegCtl* c, const LMState& S, const LMConfig& cfg) {
  double lg[6];
  pose_log(pose_mul(pose_inv(pose_from7(c->pose)), pose_from7(S.x)), lg);
  double mse = 0;
  for (int i = 0; i < 6; i++) mse += lg[i] * lg[i];
  const int before = c->outer;
  bool conv;
  if (ALGO == SICP_ALGO_SEMANTIC) conv = (mse < cfg.mse_stop) || (before + 1 > cfg.outer_cap);  // count++ first (semantic_icp.hpp:47)
  else conv = (mse < cfg.mse_stop) || (before > cfg.outer_cap);
  if (before < 64) { for (int i = 0; i < 7; i++) c->pass_pose[before][i] = S.x[i]; c->pass_lm_iters[before] = S.iter; }
  for (int i = 0; i < 7; i++) c->pose[i] = S.x[i];
  c->outer = before + 1;
  c->lm_iters_total += S.iter;
  c->lm_evals_total += S.evals;
  c->term_last = S.term;
  c->n_corr_last = c->n_corr_pass;
  c->n_corr_pass = 0;
  c->final_cost = S.cost;
  c->last_mse = mse;
  if (conv && !(mse < cfg.mse_stop)) c->flags |= 1;
  c->converged = conv ? 1 : 0;
  return conv;
}

// One inner solve (+ the outer-loop test) per launch.  Block 0 is the CONTROLLER: after every sweep the other blocks
// signal arrival on a counter and wait on a generation flag; block 0 waits for the counter, sums the block partials,
// runs the LM control step on solver state that stays in ITS shared memory for the whole solve, and publishes the next
// pose.  One atomic + one flag per LM iteration; no grid-wide barrier, no state traffic through global memory.
// ONE 256-thread CTA per SM: with up to 255 registers per thread the k_c residual chains of a slot interleave without
// spills, which feeds the FP64 pipe better than twice the warps at 128 registers (measured: 1.97 -> 1.76 ms of LM per
// KITTI EM registration), and every CTA sees the same SM so none finishes early behind an older neighbour.
template <int ALGO, int KC, int THREADS>
__device__ __forceinline__ void lm_body(const LMArgs& a) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  double* s_acc = reinterpret_cast<double*>(s_dyn);                          // [kAcc][THREADS] block_reduce staging
  char* s_pipe = reinterpret_cast<char*>(s_dyn) + sizeof(double) * kAcc * THREADS;  // [THREADS/32][2][group bytes] record double buffers
  __shared__ __align__(8) unsigned long long s_bar[2 * (THREADS / 32)];
  __shared__ LMState S;
  __shared__ double s_red[THREADS / 32][kAcc];
  __shared__ double s_tot[kAcc];
  __shared__ double s_rot[kAcc];
  __shared__ double s_x[8];  // pose to evaluate [7] + done flag
  __shared__ double s_RT[12];  // its rotation matrix (row-major) and translation
  __shared__ double s_L[36];   // Cholesky factor scratch of lm_control_warp
  __shared__ double2 s_logtab[kLogTab];  // log_tab's (c_i, -log c_i)
  fill_log_table(s_logtab, threadIdx.x, THREADS);
  LMSync* sy = a.sync;
  const bool eval_only = a.eval_pose != nullptr;
  const bool controller = blockIdx.x == 0;
  if (threadIdx.x < 7) s_x[threadIdx.x] = eval_only ? a.eval_pose[threadIdx.x] : a.ctl->pose[threadIdx.x];
  if (threadIdx.x == 7) s_x[7] = eval_only ? 0.0 : (double)a.ctl->converged;
  if (controller && threadIdx.x == 0) S.started = 0;
  GroupPipe pp;
  pp.buf = s_pipe + (size_t)(threadIdx.x >> 5) * 2 * Rec::group_bytes(KC);
  pp.bar = smem_u32(&s_bar[2 * (threadIdx.x >> 5)]);
  pp.uses = 0;
  pp.primed_src = nullptr;
  if ((threadIdx.x & 31) == 0) {
    mbar_init(pp.bar, 1);
    mbar_init(pp.bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int g0, g1;
  {
    const long long units = 8ll * (gridDim.x - 1) + a.cfg.ctl_share8;  // eighths of a share
    const long long u0 = blockIdx.x == 0 ? 0 : a.cfg.ctl_share8 + 8ll * (blockIdx.x - 1);
    group_range(a.sv.nslots >> 5, u0, blockIdx.x == 0 ? a.cfg.ctl_share8 : u0 + 8, units, &g0, &g1);
  }
  const char* my_first = first_block<KC>(a.rec, g0, g1);
  __syncthreads();
  if (s_x[7] != 0.0) {  // registration already converged: passes enqueued ahead of the host return at once
    if (a.cond && controller && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0u);  // never leave a graph loop running
    return;
  }
  unsigned gen = ld_acquire(&sy->flag);  // generations continue across launches
  long long t_comp = 0, t_ctl = 0, t_wait = 0, t_red = 0, t_lm = 0;
  const long long t_start = clock64();
  for (;;) {
    long long t0 = clock64();
    if (threadIdx.x == 0) {
      quat_to_R(s_x, s_RT);
      s_RT[9] = s_x[4]; s_RT[10] = s_x[5]; s_RT[11] = s_x[6];
    }
    __syncthreads();
    double acc[kAcc];
    sweep_acc<ALGO, KC, THREADS>(a, s_RT, s_logtab, acc, pp, g0, g1, eval_only ? nullptr : my_first);
    block_reduce<THREADS>(acc, s_acc, a.partials);
    gen++;
    __syncthreads();  // the block's partial sums are written; thread 0's gpu-scope fence below is cumulative over them
    t_comp += clock64() - t0;
    t0 = clock64();
    if (!controller) {
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&sy->count, 1u);
      }
      if (eval_only) { pipe_drain<KC, THREADS>(pp); return; }
      // lanes 0..7 spin on the generation flag (one transaction per poll; no nanosleep: its wake-up granularity is of the
      // order of a microsecond, longer than the whole control step it would be waiting for) and then fetch their word of
      // the broadcast record — flag and record share a 128-byte line
      if (threadIdx.x < 8) {
        while ((int)(ld_acquire(&sy->flag) - gen) < 0) { }
        s_x[threadIdx.x] = __ldcg(&sy->bcast[threadIdx.x]);
      }
      __syncthreads();
      t_wait += clock64() - t0;
    } else {
      if (threadIdx.x == 0) while (ld_acquire(&sy->count) != gridDim.x - 1) { }
      __syncthreads();
      t_wait += clock64() - t0;
      t0 = clock64();
      reduce_partials<THREADS>(a.partials, s_red, s_tot);
      rotate_totals(s_tot, s_RT, s_rot);
      __syncthreads();
      const long long t1 = clock64();
      t_red += t1 - t0;
      if (eval_only) {
        if (threadIdx.x < kAcc) a.eval_out[threadIdx.x] = s_rot[threadIdx.x];
        if (threadIdx.x == 0) sy->count = 0;
        pipe_drain<KC, THREADS>(pp);
        return;
      }
      if (threadIdx.x < 32) {
        const int was_started = S.started;
        __syncwarp();
        if (threadIdx.x == 0 && !was_started) {
          double* ss = reinterpret_cast<double*>(&S);
          for (int i = 0; i < (int)(sizeof(LMState) / sizeof(double)); i++) ss[i] = 0.0;
          for (int i = 0; i < 7; i++) S.x[i] = s_x[i];
        }
        __syncwarp();
        lm_control_warp(S, s_rot, a.cfg.max_iter, s_L, threadIdx.x);
      }
      if (threadIdx.x == 0) {
        t_lm += clock64() - t1;
        if (S.done) {
          const bool conv = finish_pass<ALGO>(a.ctl, S, a.cfg);
          if (a.cond) cudaGraphSetConditional(a.cond, conv ? 0u : 1u);  // graph WHILE body: run another pass?
        }
        for (int i = 0; i < 7; i++) { s_x[i] = S.cand[i]; sy->bcast[i] = S.cand[i]; }
        s_x[7] = S.done ? 1.0 : 0.0;
        sy->bcast[7] = s_x[7];
        sy->count = 0;
        st_release(&sy->flag, gen);  // release: the broadcast record and the counter reset above are visible before the flag
      }
      __syncthreads();
      t_ctl += clock64() - t0;
    }
    if (s_x[7] != 0.0) break;
  }
  pipe_drain<KC, THREADS>(pp);
#ifdef SICP_STATS
  if (threadIdx.x == 0 && blockIdx.x < 512) { atomicAdd(&g_lm_blk[blockIdx.x][0], (unsigned long long)t_comp); atomicAdd(&g_lm_blk[blockIdx.x][1], (unsigned long long)t_wait); }
#endif
  if (threadIdx.x == 0 && controller) {  // diagnostics: controller block's sweep / wait / control cycles
    a.ctl->dbg_cycles[0] += t_comp;
    a.ctl->dbg_cycles[1] += t_wait;
    a.ctl->dbg_cycles[2] += t_ctl;
    a.ctl->dbg_cycles[3] += clock64() - t_start;
    a.ctl->dbg_cycles[4] += t_red;
    a.ctl->dbg_cycles[6] += t_lm;
  }
}
template <int ALGO, int KC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) lm_kernel(LMArgs a) {
  lm_body<ALGO, KC, THREADS>(a);
}
// The same solve held to REGS registers per thread (256-thread CTA).  At ~240 registers the CTA owns the SM's whole
// register file; capped lower, the searches and E-steps of the other registrations of a batch can run beside it on
// the SM's idle issue slots (the sweep issues on ~35 % of the cycles and is bound by the FP64 pipe and its own latencies).
template <int ALGO, int KC, int REGS>
__global__ void __maxnreg__(REGS) lm_kernel_capped(LMArgs a) {
  lm_body<ALGO, KC, 256>(a);
}

// ------------------------------------------------------------------ K4 + K5 for TWO registrations at once
// The single-problem kernel idles every sweeping block while block 0 reduces, decides and publishes (ncu on a 37-block
// solve: 27 % of the stall samples sit in that gap).  Here the sweeping blocks ALTERNATE between the solves of two
// independent registrations: while the controller of problem A (block 0, which does not sweep) works on A's totals,
// everyone else sweeps problem B, and vice versa (controller of B: block 1).  As long as a control step is shorter than a
// sweep, the gap disappears from the critical path.  Each problem keeps its own LMSync / partials / RegCtl, its own
// fixed-order reduction and its own LM state, so its results are what a solve on (gridDim - 2) blocks of the
// single-problem kernel gives.  The launch ends when both solves have terminated; whichever controller finishes last sets
// the graph's WHILE condition to "some registration of the pair has not converged".
struct LMPairArgs { LMArgs p[2]; };

template <int ALGO, int KC, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) lm_pair_kernel(LMPairArgs pa) {
  extern __shared__ __align__(128) unsigned char s_dyn[];
  double* s_acc = reinterpret_cast<double*>(s_dyn);
  char* s_pipe = reinterpret_cast<char*>(s_dyn) + sizeof(double) * kAcc * THREADS;
  __shared__ __align__(8) unsigned long long s_bar[2 * (THREADS / 32)];
  __shared__ LMState S;
  __shared__ double s_red[THREADS / 32][kAcc];
  __shared__ double s_tot[kAcc];
  __shared__ double s_rot[kAcc];
  __shared__ double s_x[2][8];  // per problem: pose to evaluate [7] + done flag
  __shared__ double s_RT[12];
  __shared__ double s_L[36];
  __shared__ double2 s_logtab[kLogTab];
  fill_log_table(s_logtab, threadIdx.x, THREADS);
  const int nsweep = (int)gridDim.x - 2;
  if (threadIdx.x < 16) {
    const int p = threadIdx.x >> 3, i = threadIdx.x & 7;
    s_x[p][i] = i < 7 ? pa.p[p].ctl->pose[i] : (double)pa.p[p].ctl->converged;
  }
  GroupPipe pp;
  pp.buf = s_pipe + (size_t)(threadIdx.x >> 5) * 2 * Rec::group_bytes(KC);
  pp.bar = smem_u32(&s_bar[2 * (threadIdx.x >> 5)]);
  pp.uses = 0;
  pp.primed_src = nullptr;
  if ((threadIdx.x & 31) == 0) {
    mbar_init(pp.bar, 1);
    mbar_init(pp.bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (blockIdx.x < 2) {
    // ---------------- controller of problem p
    const int p = blockIdx.x;
    const LMArgs& a = pa.p[p];
    LMSync* sy = a.sync;
    if (s_x[p][7] == 0.0) {
      unsigned gen = ld_acquire(&sy->flag);
      if (threadIdx.x == 0) S.started = 0;
      __syncthreads();
      for (;;) {
        if (threadIdx.x == 0) {
          quat_to_R(s_x[p], s_RT);
          while (ld_acquire(&sy->count) != (unsigned)nsweep) { }
        }
        gen++;
        __syncthreads();
        reduce_partials<THREADS>(a.partials, s_red, s_tot, 2);
        rotate_totals(s_tot, s_RT, s_rot);
        __syncthreads();
        if (threadIdx.x < 32) {
          const int was_started = S.started;
          __syncwarp();
          if (threadIdx.x == 0 && !was_started) {
            double* ss = reinterpret_cast<double*>(&S);
            for (int i = 0; i < (int)(sizeof(LMState) / sizeof(double)); i++) ss[i] = 0.0;
            for (int i = 0; i < 7; i++) S.x[i] = s_x[p][i];
          }
          __syncwarp();
          lm_control_warp(S, s_rot, a.cfg.max_iter, s_L, threadIdx.x);
        }
        if (threadIdx.x == 0) {
          if (S.done) finish_pass<ALGO>(a.ctl, S, a.cfg);
          for (int i = 0; i < 7; i++) { s_x[p][i] = S.cand[i]; sy->bcast[i] = S.cand[i]; }
          s_x[p][7] = S.done ? 1.0 : 0.0;
          sy->bcast[7] = s_x[p][7];
          sy->count = 0;
          st_release(&sy->flag, gen);
        }
        __syncthreads();
        if (s_x[p][7] != 0.0) break;
      }
    }
    // whichever controller finishes last decides whether the pair needs another pass
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned* pair_done = &pa.p[0].sync->pad[0];
      const unsigned old = atomicAdd(pair_done, 1u);
      if (old == 1u) {
        __threadfence();
        const int c0 = *((volatile int*)&pa.p[0].ctl->converged), c1 = *((volatile int*)&pa.p[1].ctl->converged);
        *pair_done = 0u;
        if (pa.p[0].cond) cudaGraphSetConditional(pa.p[0].cond, (c0 && c1) ? 0u : 1u);
      }
    }
    return;
  }

  // ---------------- sweeping block: alternate between the two problems
  int g0[2], g1[2];
  const char* first[2];
  unsigned gen[2];
  bool done[2], started[2] = {false, false};
#pragma unroll
  for (int p = 0; p < 2; p++) {
    group_range(pa.p[p].sv.nslots >> 5, (long long)blockIdx.x - 2, (long long)blockIdx.x - 1, nsweep, &g0[p], &g1[p]);
    first[p] = first_block<KC>(pa.p[p].rec, g0[p], g1[p]);
    done[p] = s_x[p][7] != 0.0;
    gen[p] = ld_acquire(&pa.p[p].sync->flag);
  }
  for (;;) {
    if (done[0] && done[1]) break;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      if (done[p]) continue;
      const LMArgs& a = pa.p[p];
      LMSync* sy = a.sync;
      if (started[p]) {  // the pose of the next evaluation of problem p (published while the other problem was swept)
        if (threadIdx.x < 8) {
          while ((int)(ld_acquire(&sy->flag) - gen[p]) < 0) { }
          s_x[p][threadIdx.x] = __ldcg(&sy->bcast[threadIdx.x]);
        }
        __syncthreads();
        if (s_x[p][7] != 0.0) { done[p] = true; continue; }
      }
      started[p] = true;
      if (threadIdx.x == 0) {
        quat_to_R(s_x[p], s_RT);
        s_RT[9] = s_x[p][4]; s_RT[10] = s_x[p][5]; s_RT[11] = s_x[p][6];
      }
      __syncthreads();
      double acc[kAcc];
      // what this warp reads next: the other problem's first block while that problem is still running, else this one's again
      const char* next_src = !done[1 - p] ? first[1 - p] : first[p];
      sweep_acc<ALGO, KC, THREADS>(a, s_RT, s_logtab, acc, pp, g0[p], g1[p], next_src);
      block_reduce<THREADS>(acc, s_acc, a.partials);
      gen[p]++;
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&sy->count, 1u);
      }
    }
  }
  pipe_drain<KC, THREADS>(pp);
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
      tsl[c] = ts;
      gate[c] = (!(mahal > -700.0) && density_is_zero((two_pi * two_pi * two_pi) * det_S(nt, m, kappa), mahal)) ? 0.0 : 1.0;
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

// ------------------------------------------------------------------ pose averaging / fusion (impl/semantic_icp.hpp:169-265)
// SemanticIterativeClosestPoint::iterativeMean: Karcher mean on SE(3).  One warp; the lanes take the logarithms of
// tAverage^-1 * T_j, lane 0 adds them in index order (the reference's summation order) and steps the mean.
__global__ void iterative_mean_kernel(const double* __restrict__ poses7, int n, int max_iter, double* __restrict__ out7, int* __restrict__ converged) {
  extern __shared__ double s_log[];  // [n][6]
  __shared__ double s_avg[7];
  __shared__ int s_done;
  const int lane = threadIdx.x;
  if (lane < 7) s_avg[lane] = poses7[lane];  // tAverage = in.front()
  if (lane == 0) s_done = 0;
  __syncwarp();
  const double w = 1.0 / (double)n;
  for (int it = 0; it < max_iter; it++) {
    const Pose inv = pose_inv(pose_from7(s_avg));
    for (int j = lane; j < n; j += 32) pose_log(pose_mul(inv, pose_from7(poses7 + 7 * (size_t)j)), s_log + 6 * (size_t)j);
    __syncwarp();
    if (lane == 0) {
      double avg[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 0; j < n; j++)
        for (int c = 0; c < 6; c++) avg[c] += w * s_log[6 * (size_t)j + c];
      const Pose cur = pose_from7(s_avg);
      const Pose nxt = pose_mul(cur, pose_exp(avg));
      double d[6], sq = 0;
      pose_log(pose_mul(pose_inv(nxt), cur), d);
      for (int c = 0; c < 6; c++) sq += d[c] * d[c];
      pose_to7(nxt, s_avg);
      if (sq < 0.01) s_done = 1;  // semantic_icp.hpp:184
    }
    __syncwarp();
    if (s_done) break;
  }
  // on failure the reference returns the average of the second-to-last step... no: `tAverage = newTAverage` every
  // iteration, so the value after the last step is what it returns either way (semantic_icp.hpp:187-190)
  if (lane < 7) out7[lane] = s_avg[lane];
  if (lane == 0) *converged = s_done;
}

// SemanticIterativeClosestPoint::poseFusion: minimise  1/2 sum_n Huber_10( r_n^2 ),  r_n = e_n^T W_n e_n,
// e_n = log(T * pose_n^-1),  W_n = cov_n^-1 * scale, over T with the SE(3) local parameterisation, by the same
// Ceres-style LM controller as the registration M-step (tolerances 1e-4 * Sophus epsilon, 50,000 iterations).
// One warp: lane n owns residual block n (blocks beyond 32 are strided).  The 6-dof Jacobian of the scalar residual is
// taken by central differences of r(T * exp(h e_k)) — the reference differentiates the same function automatically.
__device__ __forceinline__ double fusion_residual(const Pose& T, const double* pinv7, const double* W) {
  double e[6];
  pose_log(pose_mul(T, pose_from7(pinv7)), e);
  double r = 0;
  for (int a = 0; a < 6; a++) {
    double s = 0;
    for (int b = 0; b < 6; b++) s += W[6 * a + b] * e[b];
    r += e[a] * s;
  }
  return r;
}
__global__ void pose_fusion_kernel(const double* __restrict__ pinv7s, const double* __restrict__ Ws, int n, const double* __restrict__ init7,
                                   int max_iter, double* __restrict__ out7, int* __restrict__ iters_out) {
  __shared__ LMState S;
  __shared__ double s_tot[kAcc];
  __shared__ double s_L[36];
  __shared__ double s_pose[7];
  const int lane = threadIdx.x;
  if (lane == 0) {
    double* ss = reinterpret_cast<double*>(&S);
    for (int i = 0; i < (int)(sizeof(LMState) / sizeof(double)); i++) ss[i] = 0.0;
    for (int i = 0; i < 7; i++) { S.x[i] = init7[i]; s_pose[i] = init7[i]; }
  }
  __syncwarp();
  const double h = 1e-6, huber_a = 10.0, huber_b = 100.0;
  for (;;) {
    const Pose T = pose_from7(s_pose);
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; i++) acc[i] = 0.0;
    for (int b = lane; b < n; b += 32) {
      const double* pi = pinv7s + 7 * (size_t)b;
      const double* W = Ws + 36 * (size_t)b;
      const double r = fusion_residual(T, pi, W);
      double j[6];
      for (int k = 0; k < 6; k++) {
        double d[6] = {0, 0, 0, 0, 0, 0};
        d[k] = h;
        const double rp = fusion_residual(pose_plus(T, d), pi, W);
        d[k] = -h;
        const double rm = fusion_residual(pose_plus(T, d), pi, W);
        j[k] = (rp - rm) / (2.0 * h);
      }
      // HuberLoss(10): rho(s) = s (s <= 100) | 20 sqrt(s) - 100; rho'' <= 0, so Ceres scales residual and Jacobian by sqrt(rho')
      const double s2 = r * r;
      const double rho0 = s2 <= huber_b ? s2 : 2.0 * huber_a * sqrt(s2) - huber_b;
      const double rho1 = s2 <= huber_b ? 1.0 : fmax(DBL_MIN, huber_a / sqrt(s2));
      int q = 0;
      for (int a = 0; a < 6; a++) {
        for (int c = 0; c <= a; c++) acc[q++] += rho1 * j[a] * j[c];
        acc[21 + a] += rho1 * j[a] * r;
      }
      acc[27] += 0.5 * rho0;
    }
#pragma unroll
    for (int i = 0; i < kAcc; i++) {  // fixed-order butterfly: deterministic
      double v = acc[i];
      for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
      if (lane == 0) s_tot[i] = v;
    }
    __syncwarp();
    lm_control_warp(S, s_tot, max_iter, s_L, lane, 0.0001 * kSophusEps, 0.0001 * kSophusEps);
    if (lane < 7) s_pose[lane] = S.cand[lane];
    __syncwarp();
    if (S.done) break;
  }
  if (lane < 7) out7[lane] = S.x[lane];
  if (lane == 0) *iters_out = S.iter;
}

// ------------------------------------------------------------------ host launchers
// Shapes of the LM kernel (LMConfig.variant).  0 is the lone-registration shape: one 256-thread CTA per SM with every
// record chain of a slot interleaved (~240 registers, the SM's whole register file).  1: 128-thread CTAs.  2-4: the
// 256-thread CTA held to 176 / 160 / 144 registers (no or a few bytes of spills): less interleave per warp, but the
// kernels of OTHER registrations become co-resident with the solve — what a batch wants (register.cu: kBatchLmVariant).
struct LmShape { int threads, minb; };
static const LmShape kLmShapes[kLmVariants] = {{256, 1}, {128, 1}, {256, 1}, {256, 1}, {256, 1}};
template <int ALGO, int KC>
static void* lm_entry_algo(int variant) {
  switch (variant) {
    case 1: return (void*)lm_kernel<ALGO, KC, 128, 1>;
    case 2: return (void*)lm_kernel_capped<ALGO, KC, 176>;
    case 3: return (void*)lm_kernel_capped<ALGO, KC, 160>;
    case 4: return (void*)lm_kernel_capped<ALGO, KC, 144>;
    default: return (void*)lm_kernel<ALGO, KC, 256, 1>;
  }
}
static void* lm_entry(int algo, int variant) {
  switch (algo) {
    case SICP_ALGO_GICP: return lm_entry_algo<SICP_ALGO_GICP, 1>(variant);
    case SICP_ALGO_SEMANTIC: return lm_entry_algo<SICP_ALGO_SEMANTIC, 1>(variant);
    default: return lm_entry_algo<SICP_ALGO_EM, 4>(variant);
  }
}
// dynamic shared memory: block_reduce staging + the per-warp double buffers of record blocks
static size_t lm_smem(int algo, int variant) {
  const int t = kLmShapes[variant].threads, kc = algo == SICP_ALGO_EM ? 4 : 1;
  return sizeof(double) * kAcc * t + (size_t)(t / 32) * 2 * Rec::group_bytes(kc);
}
// Largest cooperative grid of a shape on this device (co-resident CTAs), capped by the partials slab.
int lm_max_grid(int device, int algo, int variant) {
  static int cached[64][3][kLmVariants] = {};
  if (variant < 0 || variant >= kLmVariants) variant = 0;
  const int ai = algo == SICP_ALGO_GICP ? 0 : algo == SICP_ALGO_SEMANTIC ? 1 : 2;
  if (device >= 0 && device < 64 && cached[device][ai][variant]) return cached[device][ai][variant];
  int sms = 148, per_sm = 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  void* fn = lm_entry(algo, variant);
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm_smem(algo, variant));
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kLmShapes[variant].threads, lm_smem(algo, variant)) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int g = std::min(sms * per_sm, kLmMaxGrid);
  if (device >= 0 && device < 64) cached[device][ai][variant] = g;
  return g;
}
int lm_grid_blocks(int device) {  // lone-registration grid: one CTA of shape 0 per SM
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return std::min(sms, kLmMaxGrid);
}

sicp_status launch_estep(const sicp_cloud* src, const sicp_cloud* tgt, const LMConfig& cfg, double gate_d2, const double* d_pose7,
                         const int* d_stop, int* d_corr, const float* d_d2, char* d_rec, RegCtl* d_ctl, cudaStream_t st) {
  if (src->nslots == 0) return SICP_OK;
  const unsigned grid = (unsigned)((src->nslots + kEstepThreads - 1) / kEstepThreads);
  if (cfg.kc == 4)
    estep_kernel<4><<<grid, kEstepThreads, 0, st>>>(src->view(), tgt->view(), cfg.algo, cfg.eps, gate_d2, d_pose7, d_stop, d_corr, d_d2, d_rec, d_ctl);
  else
    estep_kernel<1><<<grid, kEstepThreads, 0, st>>>(src->view(), tgt->view(), cfg.algo, cfg.eps, gate_d2, d_pose7, d_stop, d_corr, d_d2, d_rec, d_ctl);
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}

static sicp_status launch_lm_args(LMArgs& args, int grid, cudaStream_t st) {
  const int variant = (args.cfg.variant >= 0 && args.cfg.variant < kLmVariants) ? args.cfg.variant : 0;
  int dev = 0;
  SICP_CUDA(cudaGetDevice(&dev));
  grid = std::max(1, std::min(grid, lm_max_grid(dev, args.cfg.algo, variant)));  // also sets the shared-memory attribute
  args.partials = reinterpret_cast<double*>(args.sync) + kLmSyncDoubles;
  if (grid == 1 || args.cfg.ctl_share8 < 0 || args.cfg.ctl_share8 > 8) args.cfg.ctl_share8 = 8;  // a lone block sweeps everything
  void* params[] = {&args};
  SICP_CUDA(cudaLaunchCooperativeKernel(lm_entry(args.cfg.algo, variant), dim3(grid), dim3(kLmShapes[variant].threads), params, lm_smem(args.cfg.algo, variant), st));
  count_launches(1);
  return SICP_OK;
}

// d_partials: kLmSyncDoubles doubles of LMSync (zeroed once when the workspace is created; generations continue across
// launches) followed by kLmMaxGrid * kAcc block partials.
sicp_status launch_lm(const sicp_cloud* src, const LMConfig& cfg, const char* d_rec, RegCtl* d_ctl,
                      double* d_partials, int grid, cudaStream_t st, unsigned long long cond_handle) {
  LMArgs args{src->view(), cfg, d_rec, d_ctl, nullptr, reinterpret_cast<LMSync*>(d_partials), nullptr, nullptr,
              (cudaGraphConditionalHandle)cond_handle};
  static_assert(sizeof(LMSync) <= sizeof(double) * kLmSyncDoubles, "LMSync must fit in front of the partials");
  return launch_lm_args(args, grid, st);
}

// Two registrations' inner solves in one launch (lm_pair_kernel).  Both problems must use the same algorithm.
static void* lm_pair_entry(int algo) {
  switch (algo) {
    case SICP_ALGO_GICP: return (void*)lm_pair_kernel<SICP_ALGO_GICP, 1, 256>;
    case SICP_ALGO_SEMANTIC: return (void*)lm_pair_kernel<SICP_ALGO_SEMANTIC, 1, 256>;
    default: return (void*)lm_pair_kernel<SICP_ALGO_EM, 4, 256>;
  }
}
sicp_status launch_lm_pair(const sicp_cloud* src0, const LMConfig& cfg0, const char* d_rec0, RegCtl* d_ctl0, double* d_partials0,
                           const sicp_cloud* src1, const LMConfig& cfg1, const char* d_rec1, RegCtl* d_ctl1, double* d_partials1,
                           int grid, cudaStream_t st, unsigned long long cond_handle) {
  SICP_REQUIRE(cfg0.algo == cfg1.algo, "a pair solve needs two registrations of the same algorithm");
  LMPairArgs pa;
  pa.p[0] = LMArgs{src0->view(), cfg0, d_rec0, d_ctl0, reinterpret_cast<double*>(d_partials0) + kLmSyncDoubles, reinterpret_cast<LMSync*>(d_partials0), nullptr, nullptr,
                   (cudaGraphConditionalHandle)cond_handle};
  pa.p[1] = LMArgs{src1->view(), cfg1, d_rec1, d_ctl1, reinterpret_cast<double*>(d_partials1) + kLmSyncDoubles, reinterpret_cast<LMSync*>(d_partials1), nullptr, nullptr, 0};
  int dev = 0, sms = 148;
  SICP_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  void* fn = lm_pair_entry(cfg0.algo);
  const size_t smem = lm_smem(cfg0.algo, 0);
  static bool attr_set[3] = {false, false, false};
  const int ai = cfg0.algo == SICP_ALGO_GICP ? 0 : cfg0.algo == SICP_ALGO_SEMANTIC ? 1 : 2;
  if (!attr_set[ai]) { SICP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set[ai] = true; }
  grid = std::max(3, std::min(grid, std::min(sms, kLmMaxGrid)));  // two controller blocks + at least one sweeping block
  void* params[] = {&pa};
  SICP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), params, smem, st));
  count_launches(1);
  return SICP_OK;
}

sicp_status launch_evaluate(const sicp_cloud* src, const LMConfig& cfg, const char* d_rec,
                            const double* d_pose7, double* d_out28, double* d_partials, int grid, cudaStream_t st) {
  LMArgs args{src->view(), cfg, d_rec, nullptr, nullptr, reinterpret_cast<LMSync*>(d_partials), d_pose7, d_out28, 0};
  return launch_lm_args(args, grid, st);
}

sicp_status launch_fused_labels(const sicp_cloud* src, const sicp_cloud* tgt, double eps, double gate_d2, const double* d_pose7, const int* d_corr,
                                const float* d_d2, uint32_t* d_labels_out, cudaStream_t st) {
  if (src->nslots == 0) return SICP_OK;
  fused_labels_kernel<<<(src->nslots + 127) / 128, 128, 0, st>>>(src->view(), tgt->view(), eps, gate_d2, d_pose7, d_corr, d_d2, d_labels_out);
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}

sicp_status launch_iterative_mean(const double* d_poses7, int n, int max_iter, double* d_out7, int* d_converged, cudaStream_t st) {
  iterative_mean_kernel<<<1, 32, sizeof(double) * 6 * (size_t)n, st>>>(d_poses7, n, max_iter, d_out7, d_converged);
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}
sicp_status launch_pose_fusion(const double* d_pinv7s, const double* d_Ws, int n, const double* d_init7, int max_iter, double* d_out7, int* d_iters,
                               cudaStream_t st) {
  pose_fusion_kernel<<<1, 32, 0, st>>>(d_pinv7s, d_Ws, n, d_init7, max_iter, d_out7, d_iters);
  count_launches(1);
  SICP_CUDA(cudaGetLastError());
  return SICP_OK;
}

}  // namespace sicp

extern "C" sicp_status sicp_debug_lm_blocks(unsigned long long* out1024, int reset) {
#ifdef SICP_STATS
  SICP_CUDA(cudaDeviceSynchronize());
  SICP_CUDA(cudaMemcpyFromSymbol(out1024, sicp::g_lm_blk, sizeof(unsigned long long) * 1024));
  if (reset) { static unsigned long long z[1024] = {0}; SICP_CUDA(cudaMemcpyToSymbol(sicp::g_lm_blk, z, sizeof z)); }
#else
  for (int i = 0; i < 1024; i++) out1024[i] = 0;
#endif
  return SICP_OK;
}
