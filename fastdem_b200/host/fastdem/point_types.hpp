// point_types.hpp — nanopcl::PointCloud restricted to the channels integrate() reads
// (fastdem/lib/nanoPCL/include/nanopcl/core/point_cloud.hpp:14-184, core/types.hpp:19-52):
// points as xyzw float4 (w = 1), optional intensity, optional color.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "fastdem/compat.hpp"

namespace nanopcl {

struct Color {
  uint8_t r = 0, g = 0, b = 0;
  constexpr Color() = default;
  constexpr Color(uint8_t r_, uint8_t g_, uint8_t b_) : r(r_), g(g_), b(b_) {}
};
struct Intensity {
  float val;
  explicit constexpr Intensity(float v = 0.0f) : val(v) {}
};

class PointCloud {
 public:
  size_t size() const { return xyzw_.size() / 4; }
  bool empty() const { return xyzw_.empty(); }
  void reserve(size_t n) { xyzw_.reserve(n * 4); }
  void clear() { xyzw_.clear(); intensity_.clear(); color_.clear(); cov_.clear(); }

  void add(float x, float y, float z) {  // impl/point_cloud_impl.hpp:116-119 — w = 1
    xyzw_.push_back(x); xyzw_.push_back(y); xyzw_.push_back(z); xyzw_.push_back(1.0f);
    if (use_intensity_) intensity_.push_back(0.0f);
    if (use_color_) { color_.push_back(0); color_.push_back(0); color_.push_back(0); }
    if (use_cov_) cov_.emplace_back();
  }
  void add(float x, float y, float z, Intensity i) {
    if (!use_intensity_) useIntensity();
    add(x, y, z);
    intensity_.back() = i.val;
  }
  void add(float x, float y, float z, const Color& c) {
    if (!use_color_) useColor();
    add(x, y, z);
    color_[color_.size() - 3] = c.r; color_[color_.size() - 2] = c.g; color_[color_.size() - 1] = c.b;
  }
  bool hasIntensity() const { return use_intensity_; }
  bool hasColor() const { return use_color_; }
  void useIntensity() { use_intensity_ = true; intensity_.resize(size(), 0.0f); }
  void useColor() { use_color_ = true; color_.resize(size() * 3, 0); }
  // covariance channel (point_cloud.hpp: useCovariance / covariances / covariance(i))
  bool hasCovariance() const { return use_cov_; }
  void useCovariance() { use_cov_ = true; cov_.resize(size()); }
  std::vector<Eigen::Matrix3f>& covariances() { return cov_; }
  const std::vector<Eigen::Matrix3f>& covariances() const { return cov_; }
  Eigen::Matrix3f& covariance(size_t i) { return cov_[i]; }
  const Eigen::Matrix3f& covariance(size_t i) const { return cov_[i]; }

  const float* xyzw() const { return xyzw_.data(); }
  const float* intensities() const { return use_intensity_ ? intensity_.data() : nullptr; }
  const uint8_t* colors() const { return use_color_ ? color_.data() : nullptr; }
  Eigen::Vector3f point(size_t i) const {
    return Eigen::Vector3f(xyzw_[4 * i], xyzw_[4 * i + 1], xyzw_[4 * i + 2]);
  }
  void setPointsXYZW(std::vector<float> xyzw) {
    xyzw_ = std::move(xyzw);
    if (use_cov_) cov_.resize(size());
  }

  const std::string& frameId() const { return frame_id_; }
  void setFrameId(const std::string& id) { frame_id_ = id; }
  uint64_t timestamp() const { return timestamp_ns_; }
  void setTimestamp(uint64_t ns) { timestamp_ns_ = ns; }

 private:
  std::vector<float> xyzw_;
  std::vector<float> intensity_;
  std::vector<uint8_t> color_;
  std::vector<Eigen::Matrix3f> cov_;
  std::string frame_id_;
  uint64_t timestamp_ns_ = 0;
  bool use_intensity_ = false, use_color_ = false, use_cov_ = false;
};

}  // namespace nanopcl

namespace fastdem {
using PointCloud = nanopcl::PointCloud;
using Color = nanopcl::Color;
}  // namespace fastdem
