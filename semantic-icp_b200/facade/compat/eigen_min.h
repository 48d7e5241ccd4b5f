// eigen_min.h — the few fixed-size Eigen types that appear in the reference's public signatures
// (Matrix3d / Matrix4d / Matrix4f / Matrix<double,N,N> / Matrix<double,N,1>, aligned_allocator), for builds where
// Eigen is not installed.  Column-major storage like Eigen's default, so data() is interchangeable.
// Used only when <Eigen/Core> is absent (see ../sicp_compat.h); with the real Eigen this file is not included.
#ifndef SICP_FACADE_EIGEN_MIN_H_
#define SICP_FACADE_EIGEN_MIN_H_
#include <cstddef>
#include <cmath>
#include <memory>
#include <ostream>

namespace Eigen {

template <typename T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
  Matrix() { for (int i = 0; i < R * C; i++) v_[i] = T(0); }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() {
    Matrix m;
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1);
    return m;
  }
  static Matrix Constant(T x) { Matrix m; for (int i = 0; i < R * C; i++) m.v_[i] = x; return m; }
  void setZero() { for (int i = 0; i < R * C; i++) v_[i] = T(0); }
  void setIdentity() { *this = Identity(); }
  T& operator()(int r, int c) { return v_[c * R + r]; }
  const T& operator()(int r, int c) const { return v_[c * R + r]; }
  T& operator()(int i) { return v_[i]; }              // vectors
  const T& operator()(int i) const { return v_[i]; }
  T& operator[](int i) { return v_[i]; }
  const T& operator[](int i) const { return v_[i]; }
  T* data() { return v_; }
  const T* data() const { return v_; }
  static int rows() { return R; }
  static int cols() { return C; }
  template <typename U>
  Matrix<U, R, C> cast() const {
    Matrix<U, R, C> o;
    for (int i = 0; i < R * C; i++) o.data()[i] = static_cast<U>(v_[i]);
    return o;
  }
  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> o;
    for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) o(c, r) = (*this)(r, c);
    return o;
  }
  template <int C2>
  Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& b) const {
    Matrix<T, R, C2> o;
    for (int r = 0; r < R; r++)
      for (int c = 0; c < C2; c++) {
        T s = T(0);
        for (int k = 0; k < C; k++) s += (*this)(r, k) * b(k, c);
        o(r, c) = s;
      }
    return o;
  }
  Matrix operator+(const Matrix& b) const { Matrix o; for (int i = 0; i < R * C; i++) o.v_[i] = v_[i] + b.v_[i]; return o; }
  Matrix operator-(const Matrix& b) const { Matrix o; for (int i = 0; i < R * C; i++) o.v_[i] = v_[i] - b.v_[i]; return o; }
  Matrix operator*(T s) const { Matrix o; for (int i = 0; i < R * C; i++) o.v_[i] = v_[i] * s; return o; }
  T squaredNorm() const { T s = T(0); for (int i = 0; i < R * C; i++) s += v_[i] * v_[i]; return s; }
  T norm() const { return std::sqrt(squaredNorm()); }

 private:
  T v_[R * C];
};

template <typename T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {
  for (int r = 0; r < R; r++) {
    for (int c = 0; c < C; c++) os << (c ? " " : "") << m(r, c);
    if (r + 1 < R) os << "\n";
  }
  return os;
}

typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
template <typename T>
using aligned_allocator = std::allocator<T>;

}  // namespace Eigen
#endif  // SICP_FACADE_EIGEN_MIN_H_
