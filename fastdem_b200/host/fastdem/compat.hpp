// compat.hpp — the few Eigen / nanogrid vocabulary types the FastDEM API surface uses.
// With Eigen on the include path the real types are used; otherwise minimal stand-ins with the
// same spelling for what the integrate() API needs (Isometry3d::Identity(), translation(),
// matrix().data(), Vector2d/3f element access).  Reference: fastdem/include/fastdem/fastdem.hpp,
// transform_interface.hpp.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#if __has_include(<Eigen/Geometry>)
#include <Eigen/Core>
#include <Eigen/Geometry>
#define FASTDEM_B200_HAVE_EIGEN 1
#else
#define FASTDEM_B200_HAVE_EIGEN 0
namespace Eigen {
template <typename T, int N>
struct VectorN {
  T v[N] = {};
  VectorN() = default;
  VectorN(T a, T b) { static_assert(N == 2, ""); v[0] = a; v[1] = b; }
  VectorN(T a, T b, T c) { static_assert(N == 3, ""); v[0] = a; v[1] = b; v[2] = c; }
  T& x() { return v[0]; }
  T& y() { return v[1]; }
  T& z() { static_assert(N >= 3, ""); return v[2]; }
  T x() const { return v[0]; }
  T y() const { return v[1]; }
  T z() const { static_assert(N >= 3, ""); return v[2]; }
  T& operator()(int i) { return v[i]; }
  T operator()(int i) const { return v[i]; }
  T& operator[](int i) { return v[i]; }
  T operator[](int i) const { return v[i]; }
  const T* data() const { return v; }
};
using Vector2d = VectorN<double, 2>;
using Vector3d = VectorN<double, 3>;
using Vector3f = VectorN<float, 3>;
using Vector2i = VectorN<int, 2>;

// column-major 3x3 float — what SensorModel::computeCovariance returns and the cloud's
// covariance channel stores (36 bytes, nanopcl/core/types.hpp:46-52)
struct Matrix3f {
  float m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  static Matrix3f Zero() { return Matrix3f(); }
  static Matrix3f Identity() { Matrix3f r; r.m[0] = r.m[4] = r.m[8] = 1.0f; return r; }
  float& operator()(int r, int c) { return m[c * 3 + r]; }
  float operator()(int r, int c) const { return m[c * 3 + r]; }
  const float* data() const { return m; }
  float* data() { return m; }
  Matrix3f operator*(float s) const { Matrix3f r; for (int i = 0; i < 9; ++i) r.m[i] = m[i] * s; return r; }
};
inline Matrix3f operator*(float s, const Matrix3f& a) { return a * s; }

// column-major 4x4, bottom row (0,0,0,1)
class Isometry3d {
 public:
  struct Mat4 {
    double m[16];
    const double* data() const { return m; }
    double operator()(int r, int c) const { return m[c * 4 + r]; }
    double& operator()(int r, int c) { return m[c * 4 + r]; }
  };
  struct TranslationRef {
    double* p;
    double& x() { return p[0]; }
    double& y() { return p[1]; }
    double& z() { return p[2]; }
    double operator()(int i) const { return p[i]; }
  };
  static Isometry3d Identity() {
    Isometry3d t;
    std::memset(t.m_.m, 0, sizeof(t.m_.m));
    t.m_.m[0] = t.m_.m[5] = t.m_.m[10] = t.m_.m[15] = 1.0;
    return t;
  }
  Isometry3d() { *this = IdentityInit(); }
  TranslationRef translation() { return TranslationRef{m_.m + 12}; }
  Vector3d translation() const { return Vector3d(m_.m[12], m_.m[13], m_.m[14]); }
  const Mat4& matrix() const { return m_; }
  Mat4& matrix() { return m_; }
  // rotate about Z by `angle` (what the reference's tests use: AngleAxisd(a, UnitZ))
  Isometry3d& rotateZ(double angle) {
    const double c = std::cos(angle), s = std::sin(angle);
    for (int r = 0; r < 3; ++r) {
      const double a = m_(r, 0), b = m_(r, 1);
      m_(r, 0) = a * c + b * s;
      m_(r, 1) = -a * s + b * c;
    }
    return *this;
  }

 private:
  struct IdentityTag {};
  static Isometry3d IdentityInit() {
    Isometry3d t(IdentityTag{});
    return t;
  }
  explicit Isometry3d(IdentityTag) {
    std::memset(m_.m, 0, sizeof(m_.m));
    m_.m[0] = m_.m[5] = m_.m[10] = m_.m[15] = 1.0;
  }
  Mat4 m_;
};
}  // namespace Eigen
#endif

namespace nanogrid {
struct Index {
  int v[2] = {0, 0};
  Index() = default;
  Index(int r, int c) { v[0] = r; v[1] = c; }
  int& operator()(int i) { return v[i]; }
  int operator()(int i) const { return v[i]; }
  bool operator==(const Index& o) const { return v[0] == o.v[0] && v[1] == o.v[1]; }
};
using Size = Index;
#if FASTDEM_B200_HAVE_EIGEN
using Position = Eigen::Vector2d;
using Length = Eigen::Vector2d;
#else
using Position = Eigen::Vector2d;
using Length = Eigen::Vector2d;
#endif

// host copy of one layer: rows x cols float32, column-major like Eigen::MatrixXf
class Matrix {
 public:
  Matrix() = default;
  Matrix(int rows, int cols) : rows_(rows), cols_(cols), d_(static_cast<size_t>(rows) * cols) {}
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  float& operator()(int r, int c) { return d_[static_cast<size_t>(c) * rows_ + r]; }
  float operator()(int r, int c) const { return d_[static_cast<size_t>(c) * rows_ + r]; }
  float* data() { return d_.data(); }
  const float* data() const { return d_.data(); }
  size_t size() const { return d_.size(); }

 private:
  int rows_ = 0, cols_ = 0;
  std::vector<float> d_;
};
}  // namespace nanogrid
