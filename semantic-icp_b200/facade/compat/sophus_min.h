// sophus_min.h — the slice of Sophus::SE3d the reference's public API exposes (construction from a 4x4, data(),
// matrix(), inverse(), operator*, log(), exp()), for builds where Sophus is not installed.
// Conventions follow SURVEY.md A.8: data() = [qx,qy,qz,qw,tx,ty,tz]; tangent = (upsilon, omega), translation first.
// Used only when <sophus/se3.hpp> is absent (see ../sicp_compat.h).
#ifndef SICP_FACADE_SOPHUS_MIN_H_
#define SICP_FACADE_SOPHUS_MIN_H_
#include <cmath>

namespace Sophus {

template <typename T>
struct Constants {
  static T epsilon() { return T(1e-10); }
};

template <typename T>
class SE3 {
 public:
  typedef Eigen::Matrix<T, 6, 1> Tangent;
  typedef Eigen::Matrix<T, 4, 4> Transformation;
  SE3() { p_[0] = p_[1] = p_[2] = T(0); p_[3] = T(1); p_[4] = p_[5] = p_[6] = T(0); }
  explicit SE3(const Transformation& m) {
    // rotation block -> unit quaternion (largest-pivot branch), translation column
    const T m00 = m(0, 0), m11 = m(1, 1), m22 = m(2, 2), tr = m00 + m11 + m22;
    T x, y, z, w;
    if (tr > T(0)) {
      T s = std::sqrt(tr + T(1)) * T(2);
      w = s / T(4); x = (m(2, 1) - m(1, 2)) / s; y = (m(0, 2) - m(2, 0)) / s; z = (m(1, 0) - m(0, 1)) / s;
    } else if (m00 > m11 && m00 > m22) {
      T s = std::sqrt(T(1) + m00 - m11 - m22) * T(2);
      w = (m(2, 1) - m(1, 2)) / s; x = s / T(4); y = (m(0, 1) + m(1, 0)) / s; z = (m(0, 2) + m(2, 0)) / s;
    } else if (m11 > m22) {
      T s = std::sqrt(T(1) + m11 - m00 - m22) * T(2);
      w = (m(0, 2) - m(2, 0)) / s; x = (m(0, 1) + m(1, 0)) / s; y = s / T(4); z = (m(1, 2) + m(2, 1)) / s;
    } else {
      T s = std::sqrt(T(1) + m22 - m00 - m11) * T(2);
      w = (m(1, 0) - m(0, 1)) / s; x = (m(0, 2) + m(2, 0)) / s; y = (m(1, 2) + m(2, 1)) / s; z = s / T(4);
    }
    const T n = std::sqrt(x * x + y * y + z * z + w * w);
    p_[0] = x / n; p_[1] = y / n; p_[2] = z / n; p_[3] = w / n;
    p_[4] = m(0, 3); p_[5] = m(1, 3); p_[6] = m(2, 3);
  }
  static SE3 fromData(const T* p7) { SE3 s; for (int i = 0; i < 7; i++) s.p_[i] = p7[i]; return s; }
  T* data() { return p_; }
  const T* data() const { return p_; }

  Eigen::Matrix<T, 3, 3> rotationMatrix() const {
    const T x = p_[0], y = p_[1], z = p_[2], w = p_[3];
    Eigen::Matrix<T, 3, 3> R;
    R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z);     R(0, 2) = 2 * (x * z + w * y);
    R(1, 0) = 2 * (x * y + w * z);     R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
    R(2, 0) = 2 * (x * z - w * y);     R(2, 1) = 2 * (y * z + w * x);     R(2, 2) = 1 - 2 * (x * x + y * y);
    return R;
  }
  Eigen::Matrix<T, 3, 1> translation() const { Eigen::Matrix<T, 3, 1> t; t(0) = p_[4]; t(1) = p_[5]; t(2) = p_[6]; return t; }
  Transformation matrix() const {
    Transformation m = Transformation::Identity();
    const Eigen::Matrix<T, 3, 3> R = rotationMatrix();
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) m(r, c) = R(r, c); m(r, 3) = p_[4 + r]; }
    return m;
  }
  SE3 inverse() const {
    SE3 o;
    o.p_[0] = -p_[0]; o.p_[1] = -p_[1]; o.p_[2] = -p_[2]; o.p_[3] = p_[3];
    T t[3];
    o.rotate(&p_[4], t);
    o.p_[4] = -t[0]; o.p_[5] = -t[1]; o.p_[6] = -t[2];
    return o;
  }
  SE3 operator*(const SE3& b) const {
    SE3 o;
    const T ax = p_[0], ay = p_[1], az = p_[2], aw = p_[3], bx = b.p_[0], by = b.p_[1], bz = b.p_[2], bw = b.p_[3];
    T q[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
              aw * bw - ax * bx - ay * by - az * bz};
    const T n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) o.p_[i] = q[i] / n;
    T t[3];
    rotate(&b.p_[4], t);
    for (int i = 0; i < 3; i++) o.p_[4 + i] = p_[4 + i] + t[i];
    return o;
  }
  // log: (upsilon, omega) with upsilon = V(omega)^-1 t
  Tangent log() const {
    const T x = p_[0], y = p_[1], z = p_[2], w = p_[3];
    const T n2 = x * x + y * y + z * z, n = std::sqrt(n2);
    T two_atan_over_n;
    if (n < T(1e-10)) two_atan_over_n = T(2) / w - T(2) * n2 / (T(3) * w * w * w);
    else if (std::fabs(w) < T(1e-10)) two_atan_over_n = (w > T(0) ? T(3.14159265358979323846) : -T(3.14159265358979323846)) / n;
    else two_atan_over_n = T(2) * std::atan(n / w) / n;
    const T om[3] = {two_atan_over_n * x, two_atan_over_n * y, two_atan_over_n * z};
    const T th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], th = std::sqrt(th2);
    // V^-1 = I - 1/2 Om + c Om^2,  c = (1 - th cos(th/2) / (2 sin(th/2))) / th^2
    T c;
    if (th < T(1e-5)) c = T(1) / T(12) + th2 / T(720);
    else { const T h = T(0.5) * th; c = (T(1) - th * std::cos(h) / (T(2) * std::sin(h))) / th2; }
    const T* t = &p_[4];
    const T c1[3] = {om[1] * t[2] - om[2] * t[1], om[2] * t[0] - om[0] * t[2], om[0] * t[1] - om[1] * t[0]};
    const T c2[3] = {om[1] * c1[2] - om[2] * c1[1], om[2] * c1[0] - om[0] * c1[2], om[0] * c1[1] - om[1] * c1[0]};
    Tangent o;
    for (int i = 0; i < 3; i++) { o(i) = t[i] - T(0.5) * c1[i] + c * c2[i]; o(3 + i) = om[i]; }
    return o;
  }
  static SE3 exp(const Tangent& d) {
    const T om[3] = {d(3), d(4), d(5)}, up[3] = {d(0), d(1), d(2)};
    const T th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], th = std::sqrt(th2);
    T imag, real, a, b;
    if (th < T(1e-5)) {
      imag = T(0.5) - th2 / T(48); real = T(1) - th2 / T(8); a = T(0.5) - th2 / T(24); b = T(1) / T(6) - th2 / T(120);
    } else {
      const T h = T(0.5) * th;
      imag = std::sin(h) / th; real = std::cos(h); a = (T(1) - std::cos(th)) / th2; b = (th - std::sin(th)) / (th2 * th);
    }
    SE3 o;
    o.p_[0] = imag * om[0]; o.p_[1] = imag * om[1]; o.p_[2] = imag * om[2]; o.p_[3] = real;
    const T c1[3] = {om[1] * up[2] - om[2] * up[1], om[2] * up[0] - om[0] * up[2], om[0] * up[1] - om[1] * up[0]};
    const T c2[3] = {om[1] * c1[2] - om[2] * c1[1], om[2] * c1[0] - om[0] * c1[2], om[0] * c1[1] - om[1] * c1[0]};
    for (int i = 0; i < 3; i++) o.p_[4 + i] = up[i] + a * c1[i] + b * c2[i];
    return o;
  }

 private:
  void rotate(const T* v, T* out) const {
    const Eigen::Matrix<T, 3, 3> R = rotationMatrix();
    for (int r = 0; r < 3; r++) out[r] = R(r, 0) * v[0] + R(r, 1) * v[1] + R(r, 2) * v[2];
  }
  T p_[7];
};
typedef SE3<double> SE3d;

}  // namespace Sophus
#endif  // SICP_FACADE_SOPHUS_MIN_H_
