// fdem_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A from-scratch, single-threaded C++17 restatement of the reference's
// FastDEM::integrate hot path (Ikhyeon-Cho/FastDEM @ 30d371e).  It exists only
// so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs can check and time the CUDA path against the
// reference's algorithm.  Nothing under fastdem_b200/ may include, link or call
// anything in this directory.
//
// PARITY STATUS
//   * Pinned against every known-answer value the reference's own tests assert
//     for this path (tests/test_oracle_golden.py ports them; SURVEY.md §8c list).
//   * The reference itself cannot be built here (no Eigen, yaml-cpp, gtest, and
//     nanoGrid is an un-vendored FetchContent dependency, GIT_TAG main), so there
//     is no oracle/_ref.  The grid container (nanogrid::GridMap) is restated from
//     the reference's in-tree witnesses of its convention and ANYbotics
//     grid_map_core semantics: exact index rounding at cell edges and which layers
//     move() clears are "PARITY UNPINNED" at that boundary.
//   * Floating point: every expression fixes an evaluation order (documented
//     where it is a choice); build with -ffp-contract=off so no FMA is formed.
//
// Each function cites the reference file:line it follows (paths relative to
// /root/reference/).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace fdem_oracle {

// ───────────────────────────── small fixed-size math ─────────────────────────
// Stand-ins for Eigen::Vector4f / Matrix3f / Matrix4f / Isometry3d with the
// operation order written out (SURVEY.md §8a "FP order").

struct alignas(16) Vec4f {
  float x, y, z, w;
};

struct Mat3f {  // column-major like Eigen::Matrix3f: m[c*3 + r]
  float m[9];
  float& operator()(int r, int c) { return m[c * 3 + r]; }
  float operator()(int r, int c) const { return m[c * 3 + r]; }
};

struct Mat4f {  // column-major
  float m[16];
  float operator()(int r, int c) const { return m[c * 4 + r]; }
};

struct Iso3d {  // column-major 4x4 double, bottom row (0,0,0,1)
  double m[16];
  double operator()(int r, int c) const { return m[c * 4 + r]; }
  static Iso3d identity() {
    Iso3d t{};
    t.m[0] = t.m[5] = t.m[10] = t.m[15] = 1.0;
    return t;
  }
};

// Isometry3d * Isometry3d: linear = A.lin*B.lin, translation = A.lin*B.t + A.t
// (Eigen Transform product for affine-compact-compatible modes).  3-term dots
// are evaluated left to right.
inline Iso3d compose(const Iso3d& a, const Iso3d& b) {
  Iso3d r = Iso3d::identity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      r.m[j * 4 + i] = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    }
    r.m[12 + i] =
        ((a(i, 0) * b(0, 3) + a(i, 1) * b(1, 3)) + a(i, 2) * b(2, 3)) + a(i, 3);
  }
  return r;
}

// T.matrix().cast<float>()  (nanopcl/core/transform.hpp:68-82)
inline Mat4f castf(const Iso3d& t) {
  Mat4f r;
  for (int i = 0; i < 16; ++i) r.m[i] = static_cast<float>(t.m[i]);
  return r;
}

// (A*B).rotation().cast<float>()  (fastdem/src/fastdem.cpp:182-183)
inline Mat3f rotationf(const Iso3d& t) {
  Mat3f r;
  for (int c = 0; c < 3; ++c)
    for (int row = 0; row < 3; ++row)
      r.m[c * 3 + row] = static_cast<float>(t(row, c));
  return r;
}

// Matrix4f * Vector4f as the reference's `points[i] = T * points[i]`
// (nanopcl/core/transform.hpp:26-28).  Column-accumulate order:
// r = c0*x; r = c1*y + r; r = c2*z + r; r = c3*w + r  (no FMA).
inline Vec4f mul(const Mat4f& T, const Vec4f& p) {
  float r[4];
  for (int i = 0; i < 4; ++i) {
    float acc = T(i, 0) * p.x;
    acc = T(i, 1) * p.y + acc;
    acc = T(i, 2) * p.z + acc;
    acc = T(i, 3) * p.w + acc;
    r[i] = acc;
  }
  return Vec4f{r[0], r[1], r[2], r[3]};
}

// head<3>().squaredNorm(): x² + (y² + z²)
inline float sqnorm3(float x, float y, float z) { return x * x + (y * y + z * z); }

// 3-term dot in the order a0 + (a1 + a2)
inline float dot3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

// cov = R * cov * R.transpose()  (fastdem/src/fastdem.cpp:184-187):
// tmp = R*cov evaluated first, then tmp * Rᵀ; 3-dots as a0 + (a1 + a2).
inline Mat3f rotateCov(const Mat3f& R, const Mat3f& S) {
  Mat3f tmp, out;
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k)
      tmp(i, k) = dot3(R(i, 0) * S(0, k), R(i, 1) * S(1, k), R(i, 2) * S(2, k));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      out(i, j) = dot3(tmp(i, 0) * R(j, 0), tmp(i, 1) * R(j, 1), tmp(i, 2) * R(j, 2));
  return out;
}

// ───────────────────────────── point cloud (SoA) ─────────────────────────────
// nanopcl::PointCloud restricted to the channels the path touches
// (nanopcl/core/point_cloud.hpp:14-184, core/types.hpp:19-52).

struct Color {
  uint8_t r = 0, g = 0, b = 0;
};

struct Cloud {
  std::vector<Vec4f> pts;
  std::vector<float> intensity;
  std::vector<Color> color;
  std::vector<Mat3f> cov;
  bool has_intensity = false, has_color = false, has_cov = false;

  size_t size() const { return pts.size(); }
  bool empty() const { return pts.empty(); }
  void add(float x, float y, float z) {  // impl/point_cloud_impl.hpp:116-119 (w = 1)
    pts.push_back(Vec4f{x, y, z, 1.0f});
    if (has_intensity) intensity.push_back(0.0f);
    if (has_color) color.push_back(Color{});
    if (has_cov) cov.push_back(Mat3f{});
  }
  void useCovariance() {
    has_cov = true;
    cov.assign(pts.size(), Mat3f{});
  }
  void resize(size_t n) {
    pts.resize(n);
    if (has_intensity) intensity.resize(n);
    if (has_color) color.resize(n);
    if (has_cov) cov.resize(n);
  }
};

// detail::filterInPlace (nanopcl/filters/core.hpp:22-68): stable compaction of
// every active channel.
template <typename Pred>
inline void filterInPlace(Cloud& c, Pred pred) {
  if (c.empty()) return;
  const size_t n = c.size();
  size_t w = 0;
  for (size_t r = 0; r < n; ++r) {
    if (pred(r)) {
      if (w != r) {
        c.pts[w] = c.pts[r];
        if (c.has_intensity) c.intensity[w] = c.intensity[r];
        if (c.has_color) c.color[w] = c.color[r];
        if (c.has_cov) c.cov[w] = c.cov[r];
      }
      ++w;
    }
  }
  c.resize(w);
}

// cropRange (nanopcl/filters/impl/crop_impl.hpp:79-96)
inline void cropRange(Cloud& c, float rmin, float rmax) {
  const float min_sq = rmin * rmin;
  const float max_sq = rmax * rmax;  // FLT_MAX² = +inf
  filterInPlace(c, [&](size_t i) {
    const float d2 = sqnorm3(c.pts[i].x, c.pts[i].y, c.pts[i].z);
    return d2 >= min_sq && d2 <= max_sq;
  });
}

// cropZ (crop_impl.hpp:167-178)
inline void cropZ(Cloud& c, float zmin, float zmax) {
  filterInPlace(c, [&](size_t i) {
    const float v = c.pts[i].z;
    return v >= zmin && v <= zmax;
  });
}

// transformInPlace (nanopcl/core/transform.hpp:19-37); covariances untouched.
inline void transformInPlace(Cloud& c, const Mat4f& T) {
  for (size_t i = 0; i < c.size(); ++i) c.pts[i] = mul(T, c.pts[i]);
}

// ───────────────────────────── config (mirrors fastdem::Config) ──────────────
// fastdem/include/fastdem/config/{fastdem,mapping,sensor_model,postprocess}.hpp

enum SensorType { SENSOR_CONSTANT = 0, SENSOR_LIDAR = 1, SENSOR_RGBD = 2 };
enum MappingMode { MODE_LOCAL = 0, MODE_GLOBAL = 1 };
enum EstimationType { EST_KALMAN = 0, EST_P2 = 1 };
enum MoveClearPolicy { MOVE_CLEAR_ALL = 0, MOVE_CLEAR_BASIC = 1 };

struct Config {
  // point_filter (config/fastdem.hpp:23-28)
  float z_min = -std::numeric_limits<float>::max();
  float z_max = std::numeric_limits<float>::max();
  float range_min = 0.0f;
  float range_max = std::numeric_limits<float>::max();
  // sensor_model (config/sensor_model.hpp:19-36)
  int sensor_type = SENSOR_LIDAR;
  float lidar_range_noise = 0.02f, lidar_angular_noise = 0.001f;
  float rgbd_normal_a = 0.001f, rgbd_normal_b = 0.002f, rgbd_normal_c = 0.4f,
        rgbd_lateral_factor = 0.001f;
  float constant_uncertainty = 0.03f;
  // mapping (config/mapping.hpp:8-47)
  int mode = MODE_LOCAL;
  int estimation_type = EST_KALMAN;
  float kalman_min_variance = 0.0001f, kalman_max_variance = 0.01f,
        kalman_process_noise = 0.0f;
  float p2_dn[5] = {0.01f, 0.16f, 0.50f, 0.84f, 0.99f};
  int p2_elevation_marker = 3;
  float p2_max_sample_count = 0.0f;
  // raycasting (config/postprocess.hpp:15-23)
  int raycasting_enabled = 0;
  float rc_height_conflict_threshold = 0.05f, rc_log_odds_observed = 0.4f,
        rc_log_odds_ghost = 0.2f, rc_log_odds_max = 2.0f, rc_clear_threshold = -1.0f;
  // restated-nanoGrid switch (not in the reference; see header comment)
  int move_clear_policy = MOVE_CLEAR_ALL;
};

// ───────────────────────────── sensor models ─────────────────────────────────

inline Mat3f scaledIdentity(float v) {
  Mat3f c{};
  c(0, 0) = v;
  c(1, 1) = v;
  c(2, 2) = v;
  return c;
}

// ConstantUncertaintyModel (sensors/sensor_model.hpp:87-93): variance = σ²
inline Mat3f constantCov(float uncertainty) {
  return scaledIdentity(uncertainty * uncertainty);
}

// LiDARSensorModel::computeCovariance (sensors/lidar_model.hpp:64-89).
// `cov += s * (dir * dirᵀ)` is evaluated as cov.col(j) += dir[j] * (s*dir)
// (Eigen folds the scalar into the outer product's left factor); that order is
// the oracle's definition.
inline Mat3f lidarCov(float px, float py, float pz, float range_noise, float angular_noise) {
  range_noise = std::fabs(range_noise);      // lidar_model.hpp:59-62
  angular_noise = std::fabs(angular_noise);
  const float dist_sq = sqnorm3(px, py, pz);
  if (dist_sq < 1e-6f) return scaledIdentity(0.01f);
  const float distance = std::sqrt(dist_sq);
  const float dir[3] = {px / distance, py / distance, pz / distance};
  const float var_radial = std::max(range_noise * range_noise, 1e-6f);
  const float var_lateral =
      std::max((distance * angular_noise) * (distance * angular_noise), 1e-6f);
  Mat3f cov = scaledIdentity(var_lateral);
  const float s = var_radial - var_lateral;
  const float sd[3] = {s * dir[0], s * dir[1], s * dir[2]};
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) cov(i, j) = cov(i, j) + dir[j] * sd[i];
  return cov;
}

// RGBDSensorModel::computeCovariance (sensors/rgbd_model.hpp:82-101)
inline Mat3f rgbdCov(float /*px*/, float /*py*/, float pz, float a, float b, float c, float k) {
  const float depth = pz;
  if (depth <= 0.0f) return scaledIdentity(0.01f);
  const float diff = depth - c;
  const float sigma_norm = a + b * diff * diff;  // (b*diff)*diff, left to right
  const float var_norm = sigma_norm * sigma_norm;
  const float sigma_lat = k * depth;
  const float var_lat = sigma_lat * sigma_lat;
  Mat3f cov{};
  cov(0, 0) = var_lat;
  cov(1, 1) = var_lat;
  cov(2, 2) = var_norm;
  return cov;
}

inline Mat3f sensorCov(const Config& cfg, float x, float y, float z) {
  switch (cfg.sensor_type) {
    case SENSOR_CONSTANT:
      return constantCov(cfg.constant_uncertainty);
    case SENSOR_RGBD:
      return rgbdCov(x, y, z, cfg.rgbd_normal_a, cfg.rgbd_normal_b, cfg.rgbd_normal_c,
                     cfg.rgbd_lateral_factor);
    case SENSOR_LIDAR:
    default:  // src/sensor_model.cpp:34-38 falls back to LiDAR
      return lidarCov(x, y, z, cfg.lidar_range_noise, cfg.lidar_angular_noise);
  }
}

// SensorModel::computeCovariances (sensors/sensor_model.hpp:76-85): BY-VALUE copy
// of the cloud, then one covariance per point in the sensor frame.
inline Cloud computeCovariances(const Config& cfg, Cloud scan) {
  scan.useCovariance();
  for (size_t i = 0; i < scan.size(); ++i)
    scan.cov[i] = sensorCov(cfg, scan.pts[i].x, scan.pts[i].y, scan.pts[i].z);
  return scan;
}

// FastDEM::preprocessScan (fastdem/src/fastdem.cpp:164-190)
inline Cloud preprocessScan(const Config& cfg, const Cloud& cloud, const Iso3d& T_base_sensor,
                            const Iso3d& T_world_base) {
  Cloud points = computeCovariances(cfg, cloud);
  transformInPlace(points, castf(T_base_sensor));
  cropRange(points, cfg.range_min, cfg.range_max);
  cropZ(points, cfg.z_min, cfg.z_max);
  transformInPlace(points, castf(T_world_base));
  const Mat3f R = rotationf(compose(T_world_base, T_base_sensor));
  for (size_t i = 0; i < points.size(); ++i) points.cov[i] = rotateCov(R, points.cov[i]);
  return points;
}

// ───────────────────────────── grid (restated nanoGrid) ──────────────────────
// nanogrid::GridMap is NOT in the reference tree (fastdem/CMakeLists.txt:24-28).
// Restated per SURVEY.md Appendix A from the in-tree witnesses
// (src/raycasting.cpp:63-76,112-113; include/fastdem/bridge/ros/impl.hpp:43-63;
// src/io_npz.cpp:142-144) + grid_map_core semantics.  PARITY UNPINNED.

struct Index {
  int r = 0, c = 0;
  bool operator==(const Index& o) const { return r == o.r && c == o.c; }
};
struct IndexHash {
  size_t operator()(const Index& i) const {
    return std::hash<int64_t>()((static_cast<int64_t>(i.r) << 32) ^ static_cast<uint32_t>(i.c));
  }
};

inline int wrapIndex(int i, int n) {
  if (i >= 0 && i < n) return i;
  int m = i % n;
  return m < 0 ? m + n : m;
}

namespace layer {
constexpr const char* elevation = "elevation";
constexpr const char* elevation_min = "elevation_min";
constexpr const char* elevation_max = "elevation_max";
constexpr const char* variance = "variance";
constexpr const char* n_points = "n_points";
constexpr const char* upper_bound = "upper_bound";
constexpr const char* lower_bound = "lower_bound";
constexpr const char* obstacle = "obstacle";
constexpr const char* intensity = "intensity";
constexpr const char* color = "color";
constexpr const char* kalman_p = "_kalman_p";
constexpr const char* sample_mean = "_sample_mean";
constexpr const char* sample_m2 = "_sample_m2";
constexpr const char* p2_q[5] = {"_p2_q0", "_p2_q1", "_p2_q2", "_p2_q3", "_p2_q4"};
constexpr const char* p2_n[5] = {"_p2_n0", "_p2_n1", "_p2_n2", "_p2_n3", "_p2_n4"};
constexpr const char* ghost_removal = "ghost_removal";
constexpr const char* raycasting = "raycasting";
constexpr const char* visibility_logodds = "_visibility_logodds";
constexpr const char* elevation_inpainted = "elevation_inpainted";
constexpr const char* step = "step";
constexpr const char* slope = "slope";
constexpr const char* roughness = "roughness";
constexpr const char* curvature = "curvature";
constexpr const char* normal_x = "_normal_x";
constexpr const char* normal_y = "_normal_y";
constexpr const char* normal_z = "_normal_z";
}  // namespace layer

using Matrix = std::vector<float>;  // rows*cols, column-major: [c*rows + r]

class ElevationMap {
 public:
  // ElevationMap() registers the three basic layers (elevation_map.hpp:99-103)
  ElevationMap() {
    for (const char* n : {layer::elevation, layer::elevation_min, layer::elevation_max}) {
      order_.push_back(n);
      data_[n] = Matrix();
    }
  }

  // ElevationMap::setGeometry (elevation_map.hpp:112-116): float args widened to
  // double; nanoGrid setGeometry (size = round(L/res), length = size*res,
  // position = 0, startIndex = 0) then clearAll().
  void setGeometry(float width, float height, float resolution) {
    const double L[2] = {static_cast<double>(width), static_cast<double>(height)};
    res_ = static_cast<double>(resolution);
    rows_ = static_cast<int>(std::round(L[0] / res_));
    cols_ = static_cast<int>(std::round(L[1] / res_));
    len_[0] = rows_ * res_;
    len_[1] = cols_ * res_;
    pos_[0] = pos_[1] = 0.0;
    start_ = Index{0, 0};
    for (auto& kv : data_) kv.second.assign(static_cast<size_t>(rows_) * cols_, 0.0f);
    clearAll();
  }

  int rows() const { return rows_; }
  int cols() const { return cols_; }
  double resolution() const { return res_; }
  const double* length() const { return len_; }
  const double* position() const { return pos_; }
  Index startIndex() const { return start_; }
  void setPosition(double x, double y) { pos_[0] = x; pos_[1] = y; }
  void setStartIndex(Index s) { start_ = s; }

  bool exists(const std::string& n) const { return data_.count(n) != 0; }
  void add(const std::string& n, float fill = std::numeric_limits<float>::quiet_NaN()) {
    if (!exists(n)) order_.push_back(n);
    data_[n].assign(static_cast<size_t>(rows_) * cols_, fill);
  }
  Matrix& get(const std::string& n) {
    auto it = data_.find(n);
    if (it == data_.end()) throw std::out_of_range("layer missing: " + n);
    return it->second;
  }
  const Matrix& get(const std::string& n) const {
    auto it = data_.find(n);
    if (it == data_.end()) throw std::out_of_range("layer missing: " + n);
    return it->second;
  }
  const std::vector<std::string>& layers() const { return order_; }
  size_t lin(int r, int c) const { return static_cast<size_t>(c) * rows_ + r; }
  float& at(const std::string& n, Index i) { return get(n)[lin(i.r, i.c)]; }

  void clear(const std::string& n) {
    Matrix& m = get(n);
    std::fill(m.begin(), m.end(), std::numeric_limits<float>::quiet_NaN());
  }
  void clearAll() {
    for (auto& kv : data_)
      std::fill(kv.second.begin(), kv.second.end(), std::numeric_limits<float>::quiet_NaN());
  }
  // ElevationMap::clearAt (elevation_map.hpp:131-135): every layer -> NaN
  void clearAt(Index i) {
    for (auto& kv : data_) kv.second[lin(i.r, i.c)] = std::numeric_limits<float>::quiet_NaN();
  }
  // ElevationMap::isEmpty (elevation_map.hpp:123-125)
  bool isEmpty() const {
    for (float v : get(layer::elevation))
      if (!std::isnan(v)) return false;
    return true;
  }

  // isInside / getIndex / getPosition — Appendix A
  bool isInside(double x, double y) const {
    const double tx = (pos_[0] + 0.5 * len_[0]) - x;
    const double ty = (pos_[1] + 0.5 * len_[1]) - y;
    return tx >= 0.0 && ty >= 0.0 && tx < len_[0] && ty < len_[1];
  }
  bool getIndex(double x, double y, Index& idx) const {
    const double tx = (pos_[0] + 0.5 * len_[0]) - x;
    const double ty = (pos_[1] + 0.5 * len_[1]) - y;
    if (!(tx >= 0.0 && ty >= 0.0 && tx < len_[0] && ty < len_[1])) return false;
    int ur = static_cast<int>(tx / res_);
    int uc = static_cast<int>(ty / res_);
    // tx < len can still divide to == size after rounding; grid_map rejects an
    // out-of-range index (checkIfIndexInRange)
    if (ur >= rows_ || uc >= cols_) return false;
    idx.r = wrapIndex(ur + start_.r, rows_);
    idx.c = wrapIndex(uc + start_.c, cols_);
    return true;
  }
  void getPosition(Index idx, double& x, double& y) const {
    const int ur = wrapIndex(idx.r - start_.r + rows_, rows_);
    const int uc = wrapIndex(idx.c - start_.c + cols_, cols_);
    x = pos_[0] + 0.5 * len_[0] - 0.5 * res_ - ur * res_;
    y = pos_[1] + 0.5 * len_[1] - 0.5 * res_ - uc * res_;
  }

  // move — Appendix A (grid_map_core GridMap::move).  `policy` selects which
  // layers the vacated rows/cols are reset on.
  bool move(double nx, double ny, int policy) {
    const double d[2] = {nx - pos_[0], ny - pos_[1]};
    int shift[2];   // map-frame shift in cells
    int bshift[2];  // buffer-order shift = -shift
    for (int i = 0; i < 2; ++i) {
      const double v = d[i] / res_;
      shift[i] = static_cast<int>(v + 0.5 * (v > 0 ? 1 : -1));
      bshift[i] = -shift[i];
    }
    const int size[2] = {rows_, cols_};
    int start[2] = {start_.r, start_.c};
    for (int i = 0; i < 2; ++i) {
      if (bshift[i] == 0) continue;
      if (std::abs(bshift[i]) >= size[i]) {
        clearAll();
      } else {
        const int sign = bshift[i] > 0 ? 1 : -1;
        const int st = start[i] - (sign < 0 ? 1 : 0);
        const int en = st - sign + bshift[i];
        const int n = std::abs(bshift[i]);
        int k = wrapIndex(sign > 0 ? st : en, size[i]);
        if (k + n <= size[i]) {
          clearStripe(i, k, n, policy);
        } else {
          const int first = size[i] - k;
          clearStripe(i, k, first, policy);
          clearStripe(i, 0, n - first, policy);
        }
      }
    }
    start_.r = wrapIndex(start[0] + bshift[0], rows_);
    start_.c = wrapIndex(start[1] + bshift[1], cols_);
    pos_[0] += shift[0] * res_;
    pos_[1] += shift[1] * res_;
    return bshift[0] != 0 || bshift[1] != 0;
  }

 private:
  void clearStripe(int axis, int k, int n, int policy) {
    const float nan = std::numeric_limits<float>::quiet_NaN();
    for (auto& kv : data_) {
      if (policy == MOVE_CLEAR_BASIC && kv.first != layer::elevation &&
          kv.first != layer::elevation_min && kv.first != layer::elevation_max)
        continue;
      Matrix& m = kv.second;
      if (axis == 0) {
        for (int c = 0; c < cols_; ++c)
          for (int r = k; r < k + n; ++r) m[lin(r, c)] = nan;
      } else {
        for (int c = k; c < k + n; ++c)
          for (int r = 0; r < rows_; ++r) m[lin(r, c)] = nan;
      }
    }
  }

  int rows_ = 0, cols_ = 0;
  double res_ = 0.0;
  double len_[2] = {0, 0}, pos_[2] = {0, 0};
  Index start_;
  std::vector<std::string> order_;
  std::map<std::string, Matrix> data_;
};

// colorVectorToValue (nanoGrid; pinned by tests/test_rasterization.cpp:107-127)
inline float packColor(uint8_t r, uint8_t g, uint8_t b) {
  const uint32_t bits = (static_cast<uint32_t>(r) << 16) | (static_cast<uint32_t>(g) << 8) | b;
  float v;
  std::memcpy(&v, &bits, 4);
  return v;
}

// ───────────────────────────── estimators ────────────────────────────────────

// Kalman (mapping/kalman_estimation.hpp:46-176)
class Kalman {
 public:
  Kalman(float min_var, float max_var, float q) : min_(min_var), max_(max_var), q_(q) {}
  void ensureLayers(ElevationMap& m) {  // :64-82
    if (!m.exists(layer::variance)) m.add(layer::variance, 0.0f);
    if (!m.exists(layer::n_points)) m.add(layer::n_points, 0.0f);
    if (!m.exists(layer::kalman_p)) m.add(layer::kalman_p, 0.0f);
    if (!m.exists(layer::sample_mean)) m.add(layer::sample_mean, NAN);
    if (!m.exists(layer::sample_m2)) m.add(layer::sample_m2, 0.0f);
    if (!m.exists(layer::upper_bound)) m.add(layer::upper_bound, NAN);
    if (!m.exists(layer::lower_bound)) m.add(layer::lower_bound, NAN);
  }
  void bind(ElevationMap& m) {  // :85-95
    map_ = &m;
    elev_ = &m.get(layer::elevation);
    var_ = &m.get(layer::variance);
    cnt_ = &m.get(layer::n_points);
    p_ = &m.get(layer::kalman_p);
    mean_ = &m.get(layer::sample_mean);
    m2_ = &m.get(layer::sample_m2);
    up_ = &m.get(layer::upper_bound);
    lo_ = &m.get(layer::lower_bound);
  }
  // scalar core, also used by the known-answer tests
  static void step(float& x, float& P, float& count, float& sample_mean, float& sample_var,
                   float& m2, float z, float meas_var, float min_v, float max_v, float q) {
    const float R = (meas_var > 0.0f) ? meas_var : max_v;  // :112-113
    if (std::isnan(x)) {                                   // :116-120
      x = z;
      P = R;
      count = 1.0f;
    } else {                                               // :121-127
      P += q;
      const float K = P / (P + R);
      x = x + K * (z - x);
      P = (1.0f - K) * P;
      P = std::min(std::max(P, min_v), max_v);  // std::clamp
      count += 1.0f;
    }
    if (std::isnan(sample_mean)) {                         // :130-134
      sample_mean = z;
      sample_var = 0.0f;
      m2 = 0.0f;
    } else {                                               // :135-142
      const float delta = z - sample_mean;
      const float new_mean = sample_mean + (delta / count);
      const float delta2 = z - new_mean;
      m2 += delta * delta2;
      sample_var = (count > 1.0f) ? m2 / (count - 1.0f) : 0.0f;
      sample_mean = new_mean;
    }
  }
  void update(Index idx, float z, float var) {  // :98-142
    const size_t l = map_->lin(idx.r, idx.c);
    step((*elev_)[l], (*p_)[l], (*cnt_)[l], (*mean_)[l], (*var_)[l], (*m2_)[l], z, var, min_, max_,
         q_);
  }
  void computeBounds(Index idx) {  // :145-153
    const size_t l = map_->lin(idx.r, idx.c);
    const float sigma = std::sqrt(std::max(0.0f, (*var_)[l]));
    (*up_)[l] = (*elev_)[l] + 2.0f * sigma;
    (*lo_)[l] = (*elev_)[l] - 2.0f * sigma;
  }

 private:
  float min_, max_, q_;
  ElevationMap* map_ = nullptr;
  Matrix *elev_ = nullptr, *var_ = nullptr, *cnt_ = nullptr, *p_ = nullptr, *mean_ = nullptr,
         *m2_ = nullptr, *up_ = nullptr, *lo_ = nullptr;
};

// P2Quantile (mapping/quantile_estimation.hpp:64-279)
class P2Quantile {
 public:
  P2Quantile(const float dn[5], int marker, float max_count) {  // :83-94
    marker_ = std::min(std::max(marker, 0), 4);
    max_count_ = std::max(max_count, 0.0f);
    for (int i = 0; i < 5; ++i) dn_[i] = std::min(std::max(dn[i], 0.0f), 1.0f);
    for (int i = 1; i < 5; ++i) dn_[i] = std::max(dn_[i], dn_[i - 1]);
  }
  void ensureLayers(ElevationMap& m) {  // :97-115
    if (!m.exists(layer::variance)) m.add(layer::variance, NAN);
    if (!m.exists(layer::n_points)) m.add(layer::n_points, 0.0f);
    for (int i = 0; i < 5; ++i)
      if (!m.exists(layer::p2_q[i])) m.add(layer::p2_q[i], NAN);
    for (int i = 0; i < 5; ++i)
      if (!m.exists(layer::p2_n[i])) m.add(layer::p2_n[i], static_cast<float>(i));
    if (!m.exists(layer::upper_bound)) m.add(layer::upper_bound, NAN);
    if (!m.exists(layer::lower_bound)) m.add(layer::lower_bound, NAN);
  }
  void bind(ElevationMap& m) {  // :118-138
    map_ = &m;
    elev_ = &m.get(layer::elevation);
    var_ = &m.get(layer::variance);
    cnt_ = &m.get(layer::n_points);
    for (int i = 0; i < 5; ++i) {
      q_[i] = &m.get(layer::p2_q[i]);
      n_[i] = &m.get(layer::p2_n[i]);
    }
    up_ = &m.get(layer::upper_bound);
    lo_ = &m.get(layer::lower_bound);
  }
  void update(Index idx, float x, float /*var*/) {  // :141-163
    const size_t l = map_->lin(idx.r, idx.c);
    float q[5], n[5];
    for (int k = 0; k < 5; ++k) {
      q[k] = (*q_[k])[l];
      n[k] = (*n_[k])[l];
    }
    updateP2(q, n, (*cnt_)[l], x);
    for (int k = 0; k < 5; ++k) {
      (*q_[k])[l] = q[k];
      (*n_[k])[l] = n[k];
    }
    (*elev_)[l] = ((*cnt_)[l] >= 5.0f) ? q[marker_] : x;
  }
  void computeBounds(Index idx) {  // :166-178
    const size_t l = map_->lin(idx.r, idx.c);
    (*elev_)[l] = (*q_[marker_])[l];
    const float sigma = ((*q_[3])[l] - (*q_[1])[l]) / 2.0f;
    (*var_)[l] = sigma * sigma;
    (*lo_)[l] = (*q_[0])[l];
    (*up_)[l] = (*q_[4])[l];
  }

  // :182-240
  void updateP2(float* q, float* n, float& count, float x) const {
    if (std::isnan(count) || count < 0.0f) count = 0.0f;
    if (count < 5.0f) {
      q[static_cast<int>(count)] = x;
      count += 1.0f;
      if (count >= 5.0f) {
        // std::sort on 5 floats == insertion sort (libstdc++ below 16 elements)
        for (int i = 1; i < 5; ++i) {
          const float v = q[i];
          int j = i - 1;
          while (j >= 0 && v < q[j]) {
            q[j + 1] = q[j];
            --j;
          }
          q[j + 1] = v;
        }
        for (int i = 0; i < 5; ++i) n[i] = static_cast<float>(i);
      }
      return;
    }
    int k;
    if (x < q[0]) {
      q[0] = x;
      k = 0;
    } else if (x < q[1]) {
      k = 0;
    } else if (x < q[2]) {
      k = 1;
    } else if (x < q[3]) {
      k = 2;
    } else if (x <= q[4]) {
      k = 3;
    } else {
      q[4] = x;
      k = 3;
    }
    for (int i = k + 1; i < 5; ++i) n[i] += 1.0f;
    float n_prime[5];
    for (int i = 0; i < 5; ++i) n_prime[i] = dn_[i] * count;  // pre-increment count
    count += 1.0f;
    if (max_count_ > 0.0f && count > max_count_) {
      const float scale = max_count_ / count;
      for (int i = 0; i < 5; ++i) n[i] *= scale;
      count = max_count_;
    }
    for (int i = 1; i < 4; ++i) {
      const float d = n_prime[i] - n[i];
      if ((d >= 1.0f && n[i + 1] - n[i] > 1.0f) || (d <= -1.0f && n[i - 1] - n[i] < -1.0f)) {
        const int sign = (d >= 0.0f) ? 1 : -1;
        const float q_new = parabolic(q, n, i, sign);
        q[i] = (q[i - 1] < q_new && q_new < q[i + 1]) ? q_new : linear(q, n, i, sign);
        n[i] += static_cast<float>(sign);
      }
    }
  }
  static float parabolic(const float* q, const float* n, int i, int sign) {  // :242-251
    const float d_right = n[i + 1] - n[i];
    const float d_left = n[i] - n[i - 1];
    const float d_span = n[i + 1] - n[i - 1];
    if (d_right == 0.0f || d_left == 0.0f || d_span == 0.0f) return q[i];
    const float s = static_cast<float>(sign);
    const float t1 = (d_left + s) * (q[i + 1] - q[i]) / d_right;
    const float t2 = (d_right - s) * (q[i] - q[i - 1]) / d_left;
    return q[i] + s * (t1 + t2) / d_span;
  }
  static float linear(const float* q, const float* n, int i, int sign) {  // :253-258
    const int j = i + sign;
    const float dn = n[j] - n[i];
    if (dn == 0.0f) return q[i];
    return q[i] + static_cast<float>(sign) * (q[j] - q[i]) / dn;
  }

 private:
  int marker_ = 3;
  float dn_[5];
  float max_count_ = 0.0f;
  ElevationMap* map_ = nullptr;
  Matrix *elev_ = nullptr, *var_ = nullptr, *cnt_ = nullptr, *q_[5] = {}, *n_[5] = {},
         *up_ = nullptr, *lo_ = nullptr;
};

// ───────────────────────────── ElevationMapping ──────────────────────────────
// fastdem/src/elevation_mapping.cpp (whole file), mapping/elevation_mapping.hpp

struct CellObservation {  // elevation_mapping.hpp:26-34
  float min_z = std::numeric_limits<float>::max();
  float min_z_var = 0.0f;
  float max_z = std::numeric_limits<float>::lowest();
  float max_intensity = std::numeric_limits<float>::lowest();
  float color_packed = 0.0f;
  bool has_intensity = false;
  bool has_color = false;
};
using CellObservations = std::unordered_map<Index, CellObservation, IndexHash>;

class ElevationMapping {
 public:
  ElevationMapping(ElevationMap& map, const Config& cfg)  // elevation_mapping.cpp:11-39
      : map_(map),
        cfg_(cfg),
        kalman_(cfg.kalman_min_variance, cfg.kalman_max_variance, cfg.kalman_process_noise),
        p2_(cfg.p2_dn, cfg.p2_elevation_marker, cfg.p2_max_sample_count) {
    if (cfg.estimation_type == EST_P2)
      p2_.ensureLayers(map_);
    else
      kalman_.ensureLayers(map_);
    if (!map_.exists(layer::obstacle)) map_.add(layer::obstacle, NAN);
  }

  CellObservations rasterize(const Cloud& cloud) {  // :41-92
    if (cloud.empty()) return {};
    CellObservations cells;
    cells.reserve(cloud.size());
    for (size_t i = 0; i < cloud.size(); ++i) {
      const Vec4f& pt = cloud.pts[i];
      Index index;
      if (!map_.getIndex(static_cast<double>(pt.x), static_cast<double>(pt.y), index)) continue;
      float pt_z_var = 0.0f;
      if (cloud.has_cov) pt_z_var = cloud.cov[i](2, 2);
      CellObservation& cell = cells[index];
      const float z = pt.z;
      if (z < cell.min_z) {
        cell.min_z = z;
        cell.min_z_var = pt_z_var;
      }
      if (z > cell.max_z) cell.max_z = z;
      if (cloud.has_intensity) {
        const float val = cloud.intensity[i];
        if (!cell.has_intensity || val > cell.max_intensity) {
          cell.max_intensity = val;
          cell.has_intensity = true;
        }
      }
      if (cloud.has_color) {
        cell.color_packed = packColor(cloud.color[i].r, cloud.color[i].g, cloud.color[i].b);
        cell.has_color = true;
      }
    }
    return cells;
  }

  void estimate(const CellObservations& obs) {  // :94-108
    if (obs.empty()) return;
    if (cfg_.estimation_type == EST_P2) {
      p2_.bind(map_);
      for (const auto& kv : obs) {
        p2_.update(kv.first, kv.second.min_z, kv.second.min_z_var);
        p2_.computeBounds(kv.first);
      }
    } else {
      kalman_.bind(map_);
      for (const auto& kv : obs) {
        kalman_.update(kv.first, kv.second.min_z, kv.second.min_z_var);
        kalman_.computeBounds(kv.first);
      }
    }
  }

  CellObservations update(const Cloud& cloud, double robot_x, double robot_y) {  // :110-125
    if (cfg_.mode == MODE_LOCAL) map_.move(robot_x, robot_y, cfg_.move_clear_policy);
    CellObservations obs = rasterize(cloud);
    if (obs.empty()) return obs;
    estimate(obs);
    updateMinMax(obs);
    updateObstacle(obs);
    if (cloud.has_intensity) updateIntensity(obs);
    if (cloud.has_color) updateColor(obs);
    return obs;
  }

 private:
  void updateMinMax(const CellObservations& obs) {  // :127-142
    Matrix& mn = map_.get(layer::elevation_min);
    Matrix& mx = map_.get(layer::elevation_max);
    for (const auto& kv : obs) {
      const size_t l = map_.lin(kv.first.r, kv.first.c);
      if (std::isnan(mn[l]) || kv.second.min_z < mn[l]) mn[l] = kv.second.min_z;
      if (std::isnan(mx[l]) || kv.second.max_z > mx[l]) mx[l] = kv.second.max_z;
    }
  }
  void updateObstacle(const CellObservations& obs) {  // :144-152
    map_.clear(layer::obstacle);
    Matrix& ob = map_.get(layer::obstacle);
    for (const auto& kv : obs)
      ob[map_.lin(kv.first.r, kv.first.c)] =
          (kv.second.max_z > kv.second.min_z) ? kv.second.max_z : NAN;
  }
  void updateIntensity(const CellObservations& obs) {  // :154-166
    if (!map_.exists(layer::intensity)) map_.add(layer::intensity, NAN);
    Matrix& im = map_.get(layer::intensity);
    for (const auto& kv : obs) {
      if (!kv.second.has_intensity) continue;
      float& stored = im[map_.lin(kv.first.r, kv.first.c)];
      if (std::isnan(stored) || kv.second.max_intensity > stored) stored = kv.second.max_intensity;
    }
  }
  void updateColor(const CellObservations& obs) {  // :168-175
    if (!map_.exists(layer::color)) map_.add(layer::color, NAN);
    Matrix& cm = map_.get(layer::color);
    for (const auto& kv : obs) {
      if (!kv.second.has_color) continue;
      cm[map_.lin(kv.first.r, kv.first.c)] = kv.second.color_packed;
    }
  }

  ElevationMap& map_;
  Config cfg_;
  Kalman kalman_;
  P2Quantile p2_;
};

// ───────────────────────────── voxelGrid(ANY) ────────────────────────────────
// nanopcl/core/voxel.hpp:28-42, filters/impl/voxel_grid_impl.hpp:30-236.
// The reference's std::sort on key only is unstable, so which point represents a
// multi-point voxel is implementation-defined there; the oracle DEFINES it with a
// sort on (key, index) (SURVEY.md §7 hard parts).

inline uint64_t voxelPack(float x, float y, float z, float inv) {
  constexpr int32_t OFF = 1 << 20;
  auto q = [&](float v) {
    int32_t i = static_cast<int32_t>(std::floor(v * inv));
    i = std::min(std::max(i, -OFF), OFF - 1);
    return static_cast<uint64_t>(i + OFF);
  };
  return (q(z) << 42) | (q(y) << 21) | q(x);
}

// returns the indices (into `cloud`) of the selected representatives, in voxel
// key order — exactly the order the reference pushes them into `result`.
inline std::vector<uint32_t> voxelGridAnyIndices(const Cloud& cloud, float voxel_size) {
  if (voxel_size < 0.001f || voxel_size > 100.0f)
    throw std::invalid_argument("voxel_size must be in [0.001, 100]");
  std::vector<uint32_t> out;
  if (cloud.empty()) return out;
  const float inv = 1.0f / voxel_size;
  struct KI {
    uint64_t key;
    uint32_t index;
  };
  std::vector<KI> idx;
  idx.reserve(cloud.size());
  for (size_t i = 0; i < cloud.size(); ++i) {
    const Vec4f& p = cloud.pts[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    idx.push_back({voxelPack(p.x, p.y, p.z, inv), static_cast<uint32_t>(i)});
  }
  if (idx.empty()) return out;
  std::sort(idx.begin(), idx.end(), [](const KI& a, const KI& b) {
    return a.key != b.key ? a.key < b.key : a.index < b.index;
  });
  size_t start = 0;
  while (start < idx.size()) {
    const uint64_t key = idx[start].key;
    size_t end = start + 1;
    while (end < idx.size() && idx[end].key == key) ++end;
    const size_t count = end - start;
    out.push_back(idx[start + (count * 7 + start * 13) % count].index);  // :171-172
    start = end;
  }
  return out;
}

inline Cloud extract(const Cloud& c, const std::vector<uint32_t>& sel) {
  Cloud r;
  r.has_intensity = c.has_intensity;
  r.has_color = c.has_color;
  r.has_cov = c.has_cov;
  for (uint32_t i : sel) {
    r.pts.push_back(c.pts[i]);
    if (c.has_intensity) r.intensity.push_back(c.intensity[i]);
    if (c.has_color) r.color.push_back(c.color[i]);
    if (c.has_cov) r.cov.push_back(c.cov[i]);
  }
  return r;
}

// ───────────────────────────── raycasting ────────────────────────────────────
// fastdem/src/raycasting.cpp (whole file)

namespace rc {
constexpr float kMinRayLength = 1e-4f;
constexpr float kInfinity = 1e30f;

inline void traceRay(const ElevationMap& map, float resolution, const float start[3],
                     const float end[3], Matrix& ray_min, std::vector<Index>& ray_cells) {  // :46-139
  const float dx = end[0] - start[0];
  const float dy = end[1] - start[1];
  const float ray_len_2d = std::sqrt(dx * dx + dy * dy);
  if (ray_len_2d < kMinRayLength) return;
  const float dz = end[2] - start[2];
  const int nrows = map.rows();
  const int ncols = map.cols();
  const Index buf_start = map.startIndex();
  const float origin_x = static_cast<float>(map.position()[0]) + nrows * resolution * 0.5f;
  const float origin_y = static_cast<float>(map.position()[1]) + ncols * resolution * 0.5f;
  const float gr0 = (origin_x - start[0]) / resolution;
  const float gc0 = (origin_y - start[1]) / resolution;
  const float gr1 = (origin_x - end[0]) / resolution;
  const float gc1 = (origin_y - end[1]) / resolution;
  const float dr = gr1 - gr0;
  const float dc = gc1 - gc0;
  int r = static_cast<int>(std::floor(gr0));
  int c = static_cast<int>(std::floor(gc0));
  int step_r, step_c;
  float t_max_r, t_max_c, t_delta_r, t_delta_c;
  if (std::fabs(dr) > 1e-8f) {
    step_r = (dr > 0) ? 1 : -1;
    const float boundary = (step_r > 0) ? (r + 1.0f) : static_cast<float>(r);
    t_max_r = (boundary - gr0) / dr;
    t_delta_r = static_cast<float>(step_r) / dr;
  } else {
    step_r = 0;
    t_max_r = kInfinity;
    t_delta_r = kInfinity;
  }
  if (std::fabs(dc) > 1e-8f) {
    step_c = (dc > 0) ? 1 : -1;
    const float boundary = (step_c > 0) ? (c + 1.0f) : static_cast<float>(c);
    t_max_c = (boundary - gc0) / dc;
    t_delta_c = static_cast<float>(step_c) / dc;
  } else {
    step_c = 0;
    t_max_c = kInfinity;
    t_delta_c = kInfinity;
  }
  const int max_steps = nrows + ncols;
  for (int s = 0; s < max_steps; ++s) {
    if (r >= 0 && r < nrows && c >= 0 && c < ncols) {
      const int mr = (r + buf_start.r) % nrows;
      const int mc = (c + buf_start.c) % ncols;
      const float t_exit = std::min(t_max_r, t_max_c);
      const float height = start[2] + std::min(t_exit, 1.0f) * dz;
      float& cur_min = ray_min[map.lin(mr, mc)];
      if (std::isnan(cur_min)) {
        cur_min = height;
        ray_cells.push_back(Index{mr, mc});
      } else if (height < cur_min) {
        cur_min = height;
      }
    }
    if (t_max_r < t_max_c) {
      if (t_max_r >= 1.0f) break;
      r += step_r;
      t_max_r += t_delta_r;
    } else {
      if (t_max_c >= 1.0f) break;
      c += step_c;
      t_max_c += t_delta_c;
    }
  }
}
}  // namespace rc

// applyRaycasting (raycasting.cpp:218-249) = guards + processScan (:150-179) +
// resolveGhostCells (:188-214)
inline void applyRaycasting(ElevationMap& map, const Cloud& scan, const float sensor_origin[3],
                            const Config& cfg) {
  if (!cfg.raycasting_enabled || scan.empty()) return;
  if (!map.exists(layer::elevation)) return;
  if (!map.isInside(static_cast<double>(sensor_origin[0]), static_cast<double>(sensor_origin[1])))
    return;
  if (!map.exists(layer::ghost_removal)) map.add(layer::ghost_removal);
  if (!map.exists(layer::raycasting)) map.add(layer::raycasting);
  if (!map.exists(layer::visibility_logodds)) map.add(layer::visibility_logodds);
  map.clear(layer::raycasting);

  Matrix& logodds_mat = map.get(layer::visibility_logodds);
  Matrix& min_height_mat = map.get(layer::raycasting);
  const float resolution = static_cast<float>(map.resolution());
  std::vector<Index> ray_cells;
  for (size_t i = 0; i < scan.size(); ++i) {
    const float pt[3] = {scan.pts[i].x, scan.pts[i].y, scan.pts[i].z};
    Index idx;
    if (map.getIndex(static_cast<double>(pt[0]), static_cast<double>(pt[1]), idx)) {
      float& lo = logodds_mat[map.lin(idx.r, idx.c)];
      if (std::isnan(lo)) lo = 0.0f;
      lo = std::min(lo + cfg.rc_log_odds_observed, cfg.rc_log_odds_max);
    }
    if (pt[2] >= sensor_origin[2]) continue;
    rc::traceRay(map, resolution, sensor_origin, pt, min_height_mat, ray_cells);
  }

  const Matrix& elev = map.get(layer::elevation);
  for (const Index& idx : ray_cells) {
    const size_t l = map.lin(idx.r, idx.c);
    if (std::isnan(elev[l])) continue;
    if (elev[l] > min_height_mat[l] + cfg.rc_height_conflict_threshold) {
      float& lo = logodds_mat[l];
      if (std::isnan(lo)) lo = 0.0f;
      lo -= cfg.rc_log_odds_ghost;
      if (lo < cfg.rc_clear_threshold) {
        map.clearAt(idx);
        map.at(layer::ghost_removal, idx) = 1.0f;
      }
    }
  }
}

// ───────────────────────────── inpainting ("next" row) ───────────────────────
// fastdem/src/inpainting.cpp:21-67.  neighbors() = in-bounds LOGICAL 3x3
// neighbours (Appendix A), visited row-major over the 3x3 offsets.
inline void applyInpainting(ElevationMap& map, int max_iterations, int min_valid_neighbors,
                            bool inplace) {
  const char* output = inplace ? layer::elevation : layer::elevation_inpainted;
  if (!map.exists(output)) map.add(output, NAN);
  Matrix& inpainted = map.get(output);
  if (!inplace) inpainted = map.get(layer::elevation);
  const int R = map.rows(), C = map.cols();
  const Index st = map.startIndex();
  Matrix buffer(inpainted.size());
  for (int iter = 0; iter < max_iterations; ++iter) {
    bool changed = false;
    buffer = inpainted;
    for (int lc = 0; lc < C; ++lc) {
      for (int lr = 0; lr < R; ++lr) {
        const size_t l = map.lin(wrapIndex(lr + st.r, R), wrapIndex(lc + st.c, C));
        if (!std::isnan(inpainted[l])) continue;
        float sum = 0.0f;
        int count = 0;
        for (int dr = -1; dr <= 1; ++dr) {
          for (int dc = -1; dc <= 1; ++dc) {
            if (dr == 0 && dc == 0) continue;
            const int nr = lr + dr, nc = lc + dc;
            if (nr < 0 || nr >= R || nc < 0 || nc >= C) continue;
            const float val = inpainted[map.lin(wrapIndex(nr + st.r, R), wrapIndex(nc + st.c, C))];
            if (std::isfinite(val)) {
              sum += val;
              ++count;
            }
          }
        }
        if (count >= min_valid_neighbors) {
          buffer[l] = sum / static_cast<float>(count);
          changed = true;
        }
      }
    }
    inpainted = buffer;
    if (!changed) break;
  }
}

// ───────────────────────────── spatial smoothing ("next" row) ────────────────
// applySpatialSmoothing (include/fastdem/postprocess/spatial_smoothing.hpp:38-67): median of
// the finite values in the kernel_size x kernel_size logical neighbourhood (centre included),
// in place with a double buffer; cells that are not finite or have fewer than
// min_valid_neighbors finite values keep their value.  nth_element at size/2 == the element of
// rank size/2 in sorted order.  kernel_size must be odd (the even-size region of nanoGrid is
// not pinned by anything in the reference).
inline void applySpatialSmoothing(ElevationMap& map, const std::string& layer_name,
                                  int kernel_size = 3, int min_valid_neighbors = 5) {
  if (!map.exists(layer_name)) return;
  const Matrix input = map.get(layer_name);
  Matrix& output = map.get(layer_name);
  const int R = map.rows(), C = map.cols(), h = kernel_size / 2;
  const Index st = map.startIndex();
  std::vector<float> window;
  for (int lc = 0; lc < C; ++lc) {
    for (int lr = 0; lr < R; ++lr) {
      const size_t l = map.lin(wrapIndex(lr + st.r, R), wrapIndex(lc + st.c, C));
      if (!std::isfinite(input[l])) continue;
      window.clear();
      for (int dr = -h; dr <= h; ++dr)
        for (int dc = -h; dc <= h; ++dc) {
          const int nr = lr + dr, nc = lc + dc;
          if (nr < 0 || nr >= R || nc < 0 || nc >= C) continue;
          const float val = input[map.lin(wrapIndex(nr + st.r, R), wrapIndex(nc + st.c, C))];
          if (std::isfinite(val)) window.push_back(val);
        }
      if (static_cast<int>(window.size()) < min_valid_neighbors) continue;
      const size_t mid = window.size() / 2;
      std::nth_element(window.begin(), window.begin() + mid, window.end());
      output[l] = window[mid];
    }
  }
}

// ───────────────────────────── circular regions ("next" rows) ────────────────
// nanogrid::GridMap::region(radius) / neighbors(cell, region) are NOT in the tree (nanoGrid is
// an un-vendored dependency) — PARITY UNPINNED.  Restated from the call sites: a region is a
// list of (d_row, d_col, dist_sq [m^2]) offsets; neighbors() yields the offsets whose LOGICAL
// cell is inside the map.  The 3x3 box region contains the centre (inpainting.cpp:45 skips it
// by hand), so the circular one does too.  Definition used here and in the kernels: offsets
// with dist_sq = (float)((dr^2 + dc^2) * res^2) <= radius^2, dr outer / dc inner, half-width
// ceil(radius / res).  Pinned only loosely: a 0.6 m radius on a 0.5 m grid must reach the four
// edge neighbours (test_postprocess.cpp:193-225: "slightly more than 1 cell").
struct RegionEntry {
  int dr, dc;
  float dist_sq;
};
inline std::vector<RegionEntry> circularRegion(float radius, double res) {
  std::vector<RegionEntry> out;
  const int h = static_cast<int>(std::ceil(static_cast<double>(radius) / res));
  const float r2 = radius * radius;
  for (int dr = -h; dr <= h; ++dr)
    for (int dc = -h; dc <= h; ++dc) {
      const float d2 = static_cast<float>(static_cast<double>(dr * dr + dc * dc) * res * res);
      if (d2 <= r2) out.push_back({dr, dc, d2});
    }
  return out;
}

// ───────────────────────────── uncertainty fusion ("next" row) ───────────────
// fastdem/src/uncertainty_fusion.cpp:103-186 (+ SimpleWeightedECDF :36-97).  The reference
// sorts the samples with std::sort (unstable): equal values may swap, which can only change
// the float summation order of their weights; the restatement uses a stable insertion sort.
struct UncertaintyFusionConfig {
  bool enabled = false;
  float search_radius = 0.15f, spatial_sigma = 0.05f, quantile_lower = 0.01f, quantile_upper = 0.99f;
  int min_valid_neighbors = 3;
};
struct WeightedSample {
  float value, weight;
};
inline float weightedQuantile(std::vector<WeightedSample>& s, float p) {
  if (s.empty()) return NAN;
  if (s.size() == 1) return s[0].value;
  for (size_t i = 1; i < s.size(); ++i) {  // stable insertion sort by value
    const WeightedSample v = s[i];
    size_t j = i;
    while (j > 0 && v.value < s[j - 1].value) {
      s[j] = s[j - 1];
      --j;
    }
    s[j] = v;
  }
  float total = 0.0f;
  for (const auto& x : s) total += x.weight;
  if (total <= 0.0f) return NAN;
  const float target = p * total;
  float cumulative = 0.0f;
  for (const auto& x : s) {
    cumulative += x.weight;
    if (cumulative >= target) return x.value;
  }
  return s.back().value;
}
inline void applyUncertaintyFusion(ElevationMap& map, const UncertaintyFusionConfig& cfg) {
  if (!cfg.enabled) return;
  if (!map.exists(layer::upper_bound) || !map.exists(layer::lower_bound)) return;
  Matrix& upper_mat = map.get(layer::upper_bound);
  Matrix& lower_mat = map.get(layer::lower_bound);
  const auto reg = circularRegion(cfg.search_radius, map.resolution());
  const float inv_2sigma_sq = 1.0f / (2.0f * cfg.spatial_sigma * cfg.spatial_sigma);
  Matrix upper_buffer = upper_mat, lower_buffer = lower_mat;
  const int R = map.rows(), C = map.cols();
  const Index st = map.startIndex();
  std::vector<WeightedSample> lo, up;
  for (int lc = 0; lc < C; ++lc) {
    for (int lr = 0; lr < R; ++lr) {
      const size_t l = map.lin(wrapIndex(lr + st.r, R), wrapIndex(lc + st.c, C));
      if (!std::isfinite(upper_mat[l]) || !std::isfinite(lower_mat[l])) continue;
      lo.clear();
      up.clear();
      int valid = 0;
      for (const auto& e : reg) {
        const int nr = lr + e.dr, nc = lc + e.dc;
        if (nr < 0 || nr >= R || nc < 0 || nc >= C) continue;
        const size_t nl = map.lin(wrapIndex(nr + st.r, R), wrapIndex(nc + st.c, C));
        const float nu = upper_mat[nl], nlo = lower_mat[nl];
        if (!std::isfinite(nu) || !std::isfinite(nlo)) continue;
        const float w_spatial = std::exp(-e.dist_sq * inv_2sigma_sq);
        const float range = nu - nlo;
        const float w_range = 1.0f / (range + 1e-4f);
        const float w = w_spatial * w_range;
        if (w > 1e-6f) {  // SimpleWeightedECDF::add (values are finite here)
          lo.push_back({nlo, w});
          up.push_back({nu, w});
        }
        ++valid;
      }
      if (valid >= cfg.min_valid_neighbors) {
        const float lower = weightedQuantile(lo, cfg.quantile_lower);
        const float upper = weightedQuantile(up, cfg.quantile_upper);
        if (std::isfinite(lower) && std::isfinite(upper)) {
          upper_buffer[l] = upper;
          lower_buffer[l] = lower;
        }
      }
    }
  }
  upper_mat = upper_buffer;
  lower_mat = lower_buffer;
}

// ───────────────────────────── feature extraction ("next" row) ───────────────
// fastdem/src/feature_extraction.cpp:28-118 + nanopcl::geometry::computePCA
// (nanopcl/geometry/impl/pca.hpp:67-90), which calls Eigen's
// SelfAdjointEigenSolver<Matrix3f>::computeDirect.  Eigen is not in this container: the
// closed-form 3x3 solver below restates Eigen 3.4's published algorithm (shift by trace/3,
// scale to [-1,1], trigonometric roots, eigenvectors as row cross products) — PARITY UNPINNED
// beyond the reference's own assertions (flat plane, tilted plane, step edge:
// test_postprocess.cpp:273-345).  Matrices are symmetric 3x3, stored m[i][j].
struct Eig3 {
  float val[3];     // ascending
  float vec[3][3];  // vec[k] = eigenvector of val[k]
};
inline void eig3_cross(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
inline float eig3_sqnorm(const float* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
// kernel of a rank-2 symmetric matrix (Eigen's extract_kernel)
inline void eig3_kernel(const float m[3][3], float* res, float* representative) {
  int i0 = 0;
  float best = std::fabs(m[0][0]);
  for (int i = 1; i < 3; ++i)
    if (std::fabs(m[i][i]) > best) { best = std::fabs(m[i][i]); i0 = i; }
  float col[3][3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) col[j][i] = m[i][j];
  for (int i = 0; i < 3; ++i) representative[i] = col[i0][i];
  float c0[3], c1[3];
  eig3_cross(representative, col[(i0 + 1) % 3], c0);
  eig3_cross(representative, col[(i0 + 2) % 3], c1);
  const float n0 = eig3_sqnorm(c0), n1 = eig3_sqnorm(c1);
  if (n0 > n1) { const float s = std::sqrt(n0); for (int i = 0; i < 3; ++i) res[i] = c0[i] / s; }
  else { const float s = std::sqrt(n1); for (int i = 0; i < 3; ++i) res[i] = c1[i] / s; }
}
inline Eig3 eig3_direct(const float cov[3][3]) {
  Eig3 out;
  const float eps = std::numeric_limits<float>::epsilon();
  const float shift = (cov[0][0] + cov[1][1] + cov[2][2]) / 3.0f;
  float m[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m[i][j] = (i >= j) ? cov[i][j] : cov[j][i];  // lower triangle view
  for (int i = 0; i < 3; ++i) m[i][i] -= shift;
  float scale = 0.0f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(m[i][j]));
  if (scale > 0.0f)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) m[i][j] /= scale;
  // computeRoots: x^3 - c2 x^2 + c1 x - c0 = 0
  const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[1][0] * m[2][0] * m[2][1] -
                   m[0][0] * m[2][1] * m[2][1] - m[1][1] * m[2][0] * m[2][0] - m[2][2] * m[1][0] * m[1][0];
  const float c1 = m[0][0] * m[1][1] - m[1][0] * m[1][0] + m[0][0] * m[2][2] - m[2][0] * m[2][0] +
                   m[1][1] * m[2][2] - m[2][1] * m[2][1];
  const float c2 = m[0][0] + m[1][1] + m[2][2];
  const float inv3 = 1.0f / 3.0f, sqrt3 = std::sqrt(3.0f);
  const float c2_3 = c2 * inv3;
  float a_3 = (c2 * c2_3 - c1) * inv3;
  a_3 = std::max(a_3, 0.0f);
  const float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
  float q = a_3 * a_3 * a_3 - half_b * half_b;
  q = std::max(q, 0.0f);
  const float rho = std::sqrt(a_3);
  const float theta = std::atan2(std::sqrt(q), half_b) * inv3;
  const float ct = std::cos(theta), sn = std::sin(theta);
  float ev[3] = {c2_3 - rho * (ct + sqrt3 * sn), c2_3 - rho * (ct - sqrt3 * sn), c2_3 + 2.0f * rho * ct};
  float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};  // V[k] = eigenvector k
  if (!((ev[2] - ev[0]) <= eps)) {
    float d0 = ev[2] - ev[1];
    const float d1 = ev[1] - ev[0];
    int k = 0, l = 2;
    if (d0 > d1) { k = 2; l = 0; d0 = d1; }  // Eigen: swap(k, l); d0 = d1;
    float tmp[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) tmp[i][j] = m[i][j];
    for (int i = 0; i < 3; ++i) tmp[i][i] -= ev[k];
    eig3_kernel(tmp, V[k], V[l]);
    if (d0 <= 2.0f * eps * d1) {
      const float dot = V[k][0] * V[l][0] + V[k][1] * V[l][1] + V[k][2] * V[l][2];
      for (int i = 0; i < 3; ++i) V[l][i] -= dot * V[l][i];
      const float nrm = std::sqrt(eig3_sqnorm(V[l]));
      for (int i = 0; i < 3; ++i) V[l][i] /= nrm;
    } else {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) tmp[i][j] = m[i][j];
      for (int i = 0; i < 3; ++i) tmp[i][i] -= ev[l];
      float dummy[3];
      eig3_kernel(tmp, V[l], dummy);
    }
    float c[3];
    eig3_cross(V[2], V[0], c);
    const float nrm = std::sqrt(eig3_sqnorm(c));
    for (int i = 0; i < 3; ++i) V[1][i] = c[i] / nrm;
  }
  for (int k = 0; k < 3; ++k) {
    out.val[k] = ev[k] * scale + shift;
    for (int i = 0; i < 3; ++i) out.vec[k][i] = V[k][i];
  }
  return out;
}

inline void applyFeatureExtraction(ElevationMap& map, float analysis_radius = 0.3f,
                                   int min_valid_neighbors = 4, float step_lower_percentile = 0.05f,
                                   float step_upper_percentile = 0.95f) {
  if (!map.exists(layer::elevation)) return;
  for (const char* n : {layer::step, layer::slope, layer::roughness, layer::curvature, layer::normal_x,
                        layer::normal_y, layer::normal_z})
    if (!map.exists(n)) map.add(n, NAN);
  const Matrix& elev = map.get(layer::elevation);
  Matrix& step_mat = map.get(layer::step);
  Matrix& slope_mat = map.get(layer::slope);
  Matrix& rough_mat = map.get(layer::roughness);
  Matrix& curv_mat = map.get(layer::curvature);
  Matrix& nx_mat = map.get(layer::normal_x);
  Matrix& ny_mat = map.get(layer::normal_y);
  Matrix& nz_mat = map.get(layer::normal_z);
  const double res = map.resolution();
  const auto reg = circularRegion(analysis_radius, res);
  const int R = map.rows(), C = map.cols();
  const Index st = map.startIndex();
  std::vector<float> z_vals;
  for (int lc = 0; lc < C; ++lc) {
    for (int lr = 0; lr < R; ++lr) {
      const size_t l = map.lin(wrapIndex(lr + st.r, R), wrapIndex(lc + st.c, C));
      const float center_z = elev[l];
      if (!std::isfinite(center_z)) continue;
      float sum[3] = {0, 0, 0};
      float sq[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      z_vals.clear();
      int count = 0;
      for (const auto& e : reg) {
        const int nr = lr + e.dr, nc = lc + e.dc;
        if (nr < 0 || nr >= R || nc < 0 || nc >= C) continue;
        const float nz = elev[map.lin(wrapIndex(nr + st.r, R), wrapIndex(nc + st.c, C))];
        if (!std::isfinite(nz)) continue;
        const float d[3] = {-e.dr * static_cast<float>(res), -e.dc * static_cast<float>(res), nz - center_z};
        for (int i = 0; i < 3; ++i) sum[i] += d[i];
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) sq[i][j] += d[i] * d[j];
        z_vals.push_back(nz);
        ++count;
      }
      if (count < min_valid_neighbors) continue;
      const float inv_n = 1.0f / static_cast<float>(count);
      float mean[3], cov[3][3];
      for (int i = 0; i < 3; ++i) mean[i] = sum[i] * inv_n;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov[i][j] = sq[i][j] * inv_n - mean[i] * mean[j];
      const float trace = cov[0][0] + cov[1][1] + cov[2][2];
      if (trace < std::numeric_limits<float>::epsilon()) continue;  // computePCA: valid = false
      const Eig3 pca = eig3_direct(cov);
      if (pca.val[1] < 1e-8f) continue;
      float normal[3] = {pca.vec[0][0], pca.vec[0][1], pca.vec[0][2]};
      if (normal[2] < 0.0f)
        for (float& v : normal) v = -v;
      std::sort(z_vals.begin(), z_vals.end());
      const int lo = static_cast<int>(step_lower_percentile * (count - 1));
      const int hi = static_cast<int>(step_upper_percentile * (count - 1));
      step_mat[l] = z_vals[hi] - z_vals[lo];
      slope_mat[l] = std::acos(std::fabs(normal[2])) * 180.0f / static_cast<float>(M_PI);
      rough_mat[l] = std::sqrt(pca.val[0]);
      curv_mat[l] = (trace > 0.0f) ? std::fabs(pca.val[0] / trace) : 0.0f;
      nx_mat[l] = normal[0];
      ny_mat[l] = normal[1];
      nz_mat[l] = normal[2];
    }
  }
}

// ───────────────────────────── map -> PointCloud2 ("next" row) ───────────────
// toPointCloud2Impl (fastdem/include/fastdem/bridge/ros/impl.hpp:29-174): one point per cell
// with a finite elevation, fields x, y, z, then every visible (non-'_') layer except the
// elevation layer and color as FLOAT32, then "rgb" (the packed colour bits) when the map
// has a colour layer; points in sub-region order, columns outer, rows inner, both starting at
// sub_start and wrapping around the circular buffer.
struct PackedCloud {
  std::vector<std::string> fields;  // 4 bytes each, offset = 4 * index
  uint32_t point_step = 0;
  uint32_t width = 0;
  std::vector<uint8_t> data;
};
inline PackedCloud toPointCloud2(const ElevationMap& map, const char* elevation_layer, Index sub_start,
                                 int sub_rows, int sub_cols) {
  PackedCloud msg;
  const Matrix& elev = map.get(elevation_layer);
  const int rows = map.rows(), cols = map.cols();
  const Index st = map.startIndex();
  const double res = map.resolution();
  const double origin_x = map.position()[0] + map.length()[0] / 2.0 - res / 2.0;
  const double origin_y = map.position()[1] + map.length()[1] / 2.0 - res / 2.0;
  std::vector<float> row_x(sub_rows), col_y(sub_cols);
  std::vector<int> buf_row(sub_rows), buf_col(sub_cols);
  for (int i = 0; i < sub_rows; ++i) {
    const int r = (sub_start.r + i) % rows;
    buf_row[i] = r;
    const int unwrapped = (r - st.r + rows) % rows;
    row_x[i] = static_cast<float>(origin_x - unwrapped * res);
  }
  for (int j = 0; j < sub_cols; ++j) {
    const int c = (sub_start.c + j) % cols;
    buf_col[j] = c;
    const int unwrapped = (c - st.c + cols) % cols;
    col_y[j] = static_cast<float>(origin_y - unwrapped * res);
  }
  std::vector<std::string> float_layers;
  bool has_color = false;
  for (const auto& l : map.layers()) {
    if (!l.empty() && l[0] == '_') continue;  // layer::isInternal
    if (l == elevation_layer) continue;
    if (l == layer::color) { has_color = true; continue; }
    float_layers.push_back(l);
  }
  msg.fields = {"x", "y", "z"};
  for (const auto& l : float_layers) msg.fields.push_back(l);
  if (has_color) msg.fields.push_back("rgb");
  msg.point_step = static_cast<uint32_t>(4 * msg.fields.size());
  std::vector<const float*> ptrs;
  for (const auto& l : float_layers) ptrs.push_back(map.get(l).data());
  const float* color = has_color ? map.get(layer::color).data() : nullptr;
  for (int j = 0; j < sub_cols; ++j) {
    const size_t base = static_cast<size_t>(buf_col[j]) * rows;
    for (int i = 0; i < sub_rows; ++i) {
      const size_t idx = base + buf_row[i];
      const float z = elev[idx];
      if (!std::isfinite(z)) continue;
      auto put = [&](float v) {
        uint8_t b[4];
        std::memcpy(b, &v, 4);
        msg.data.insert(msg.data.end(), b, b + 4);
      };
      put(row_x[i]);
      put(col_y[j]);
      put(z);
      for (const float* p : ptrs) put(p[idx]);
      if (color) put(color[idx]);
      ++msg.width;
    }
  }
  return msg;
}

// ───────────────────────────── FastDEM pipeline ──────────────────────────────
// fastdem/src/fastdem.cpp:122-162

struct ScanStats {
  int64_t n_input = 0;
  int64_t n_kept = 0;    // points surviving cropRange + cropZ
  int64_t n_cells = 0;   // cells touched by rasterize
  int64_t n_voxels = 0;  // ray_scan size (raycasting only)
  int32_t integrated = 0;  // integrate()'s bool
};

class FastDEM {
 public:
  FastDEM(ElevationMap& map, const Config& cfg) : map_(map), cfg_(cfg), mapping_(map, cfg) {}

  bool integrate(const Cloud& cloud, const Iso3d& T_base_sensor, const Iso3d& T_world_base,
                 ScanStats* stats = nullptr, Cloud* preprocessed_out = nullptr) {
    ScanStats local;
    ScanStats& s = stats ? *stats : local;
    s = ScanStats{};
    s.n_input = static_cast<int64_t>(cloud.size());
    if (cloud.empty()) return false;  // :125-128
    Cloud points = preprocessScan(cfg_, cloud, T_base_sensor, T_world_base);  // :135
    s.n_kept = static_cast<int64_t>(points.size());
    if (points.empty()) return false;  // :137-138
    if (preprocessed_out) *preprocessed_out = points;
    CellObservations obs = mapping_.update(points, T_world_base.m[12], T_world_base.m[13]);  // :144-145
    s.n_cells = static_cast<int64_t>(obs.size());
    if (cfg_.raycasting_enabled) {  // :153-159
      const Iso3d T = compose(T_world_base, T_base_sensor);
      const float origin[3] = {static_cast<float>(T.m[12]), static_cast<float>(T.m[13]),
                               static_cast<float>(T.m[14])};
      Cloud ray_scan =
          extract(points, voxelGridAnyIndices(points, static_cast<float>(map_.resolution())));
      s.n_voxels = static_cast<int64_t>(ray_scan.size());
      applyRaycasting(map_, ray_scan, origin, cfg_);
    }
    s.integrated = 1;
    return true;
  }

  ElevationMapping& mapping() { return mapping_; }

 private:
  ElevationMap& map_;
  Config cfg_;
  ElevationMapping mapping_;
};

}  // namespace fdem_oracle
