// se3.cuh — SE(3) arithmetic with Sophus conventions (pose7 = [qx,qy,qz,qw,tx,ty,tz], tangent = (upsilon, omega)),
// used by the device-side LM update (replaces local_parameterization_se3.h:17-36 and the Sophus calls at
// impl/gicp.hpp:153, impl/em_icp.hpp:179, impl/semantic_icp.hpp:151).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace sicp {

#ifdef __CUDA_ARCH__
#define SICP_MUL(a, b) __dmul_rn((a), (b))
#define SICP_ADD(a, b) __dadd_rn((a), (b))
#define SICP_SUB(a, b) __dsub_rn((a), (b))
#else
#define SICP_MUL(a, b) ((a) * (b))
#define SICP_ADD(a, b) ((a) + (b))
#define SICP_SUB(a, b) ((a) - (b))
#endif

struct Pose {
  double q[4];  // x,y,z,w
  double t[3];
};

__host__ __device__ inline Pose pose_from7(const double* p) {
  Pose T;
  for (int i = 0; i < 4; i++) T.q[i] = p[i];
  for (int i = 0; i < 3; i++) T.t[i] = p[4 + i];
  return T;
}
__host__ __device__ inline void pose_to7(const Pose& T, double* p) {
  for (int i = 0; i < 4; i++) p[i] = T.q[i];
  for (int i = 0; i < 3; i++) p[4 + i] = T.t[i];
}

// Eigen::Quaternion::toRotationMatrix, evaluated without FMA contraction so that the f64->f32 source transform
// (pcl::transformPointCloud<PointT,double>, impl/gicp.hpp:55-59) rounds exactly like the reference.  R row-major.
__host__ __device__ inline void quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = SICP_MUL(2.0, x), ty = SICP_MUL(2.0, y), tz = SICP_MUL(2.0, z);
  const double twx = SICP_MUL(tx, w), twy = SICP_MUL(ty, w), twz = SICP_MUL(tz, w);
  const double txx = SICP_MUL(tx, x), txy = SICP_MUL(ty, x), txz = SICP_MUL(tz, x);
  const double tyy = SICP_MUL(ty, y), tyz = SICP_MUL(tz, y), tzz = SICP_MUL(tz, z);
  R[0] = SICP_SUB(1.0, SICP_ADD(tyy, tzz)); R[1] = SICP_SUB(txy, twz);               R[2] = SICP_ADD(txz, twy);
  R[3] = SICP_ADD(txy, twz);               R[4] = SICP_SUB(1.0, SICP_ADD(txx, tzz)); R[5] = SICP_SUB(tyz, twx);
  R[6] = SICP_SUB(txz, twy);               R[7] = SICP_ADD(tyz, twx);               R[8] = SICP_SUB(1.0, SICP_ADD(txx, tyy));
}

__host__ __device__ inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__host__ __device__ inline void quat_rot(const double* q, const double* p, double* o) {
  double uv[3], c2[3];
  cross3(q, p, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  cross3(q, uv, c2);
  for (int i = 0; i < 3; i++) o[i] = p[i] + q[3] * uv[i] + c2[i];
}
__host__ __device__ inline Pose pose_mul(const Pose& A, const Pose& B) {
  Pose C;
  const double ax = A.q[0], ay = A.q[1], az = A.q[2], aw = A.q[3], bx = B.q[0], by = B.q[1], bz = B.q[2], bw = B.q[3];
  C.q[3] = aw * bw - ax * bx - ay * by - az * bz;
  C.q[0] = aw * bx + ax * bw + ay * bz - az * by;
  C.q[1] = aw * by + ay * bw + az * bx - ax * bz;
  C.q[2] = aw * bz + az * bw + ax * by - ay * bx;
  double n = sqrt(C.q[0] * C.q[0] + C.q[1] * C.q[1] + C.q[2] * C.q[2] + C.q[3] * C.q[3]);
  for (int i = 0; i < 4; i++) C.q[i] /= n;
  double rb[3];
  quat_rot(A.q, B.t, rb);
  for (int i = 0; i < 3; i++) C.t[i] = A.t[i] + rb[i];
  return C;
}
__host__ __device__ inline Pose pose_inv(const Pose& A) {
  Pose C;
  C.q[0] = -A.q[0]; C.q[1] = -A.q[1]; C.q[2] = -A.q[2]; C.q[3] = A.q[3];
  double mt[3] = {-A.t[0], -A.t[1], -A.t[2]};
  quat_rot(C.q, mt, C.t);
  return C;
}
__host__ __device__ inline void hat_sq(const double* w, double* O, double* O2) {  // row-major 3x3
  O[0] = 0; O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2]; O[4] = 0; O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += O[3 * i + k] * O[3 * k + j];
      O2[3 * i + j] = s;
    }
}
constexpr double kSophusEps = 1e-10;  // Sophus::Constants<double>::epsilon()

__host__ __device__ inline Pose pose_exp(const double* d) {
  const double* up = d;
  const double* om = d + 3;
  Pose T;
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double theta = sqrt(th2);
  double imag, real;
  if (th2 < kSophusEps * kSophusEps) {
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  T.q[0] = imag * om[0]; T.q[1] = imag * om[1]; T.q[2] = imag * om[2]; T.q[3] = real;
  double V[9];
  if (theta < kSophusEps) {
    quat_to_R(T.q, V);
  } else {
    double O[9], O2[9];
    hat_sq(om, O, O2);
    const double a = (1.0 - cos(theta)) / th2;
    const double b = (theta - sin(theta)) / (th2 * theta);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
  }
  for (int i = 0; i < 3; i++) T.t[i] = V[3 * i] * up[0] + V[3 * i + 1] * up[1] + V[3 * i + 2] * up[2];
  return T;
}
__host__ __device__ inline void pose_log(const Pose& T, double* out) {
  const double* q = T.q;
  const double sq_n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  const double w = q[3];
  double two_atan, theta;
  if (sq_n < kSophusEps * kSophusEps) {
    two_atan = 2.0 / w - (2.0 / 3.0) * sq_n / (w * w * w);
    theta = 2.0 * sq_n / w;
  } else {
    const double n = sqrt(sq_n);
    const double at = (w < 0) ? atan2(-n, -w) : atan2(n, w);
    two_atan = 2.0 * at / n;
    theta = two_atan * n;
  }
  double om[3] = {two_atan * q[0], two_atan * q[1], two_atan * q[2]};
  double O[9], O2[9], Vi[9];
  hat_sq(om, O, O2);
  double c;
  if (fabs(theta) < kSophusEps) c = 1.0 / 12.0;
  else {
    const double half = 0.5 * theta;
    c = (1.0 - theta * cos(half) / (2.0 * sin(half))) / (theta * theta);
  }
  for (int i = 0; i < 9; i++) Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + c * O2[i];
  for (int i = 0; i < 3; i++) out[i] = Vi[3 * i] * T.t[0] + Vi[3 * i + 1] * T.t[1] + Vi[3 * i + 2] * T.t[2];
  out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}
__host__ __device__ inline Pose pose_plus(const Pose& T, const double* delta) { return pose_mul(T, pose_exp(delta)); }

}  // namespace sicp
