// fastdem.hpp — fastdem::FastDEM (fastdem/include/fastdem/fastdem.hpp:55-156): same
// constructors, fluent setters, integrate() overloads, callbacks and return/error behaviour
// (fastdem/src/fastdem.cpp), forwarding to libfastdem_b200.so.  The reference's spdlog messages
// become stderr lines.
#pragma once

#include <cstdio>
#include <functional>
#include <memory>
#include <optional>
#include <string>

#include "fastdem/config.hpp"
#include "fastdem/elevation_map.hpp"
#include "fastdem/point_types.hpp"

namespace fastdem {

// fastdem/include/fastdem/transform_interface.hpp
class Calibration {
 public:
  using Ptr = std::shared_ptr<Calibration>;
  virtual ~Calibration() = default;
  virtual std::optional<Eigen::Isometry3d> getExtrinsic(const std::string& sensor_frame) const = 0;
  virtual std::string getBaseFrame() const = 0;
};
class Odometry {
 public:
  using Ptr = std::shared_ptr<Odometry>;
  virtual ~Odometry() = default;
  virtual std::optional<Eigen::Isometry3d> getPoseAt(uint64_t timestamp_ns) const = 0;
  virtual std::string getWorldFrame() const = 0;
};

class FastDEM {
 public:
  using CloudCallback = std::function<void(const PointCloud&)>;

  explicit FastDEM(ElevationMap& map) : FastDEM(map, Config{}) {}
  FastDEM(ElevationMap& map, const Config& cfg) : map_(map), cfg_(cfg) {
    fdem_config a = toAbi(cfg_);
    check(fdem_mapper_create(map_.handle(), &a, &h_));
  }
  ~FastDEM() { if (h_) fdem_mapper_destroy(h_); }
  FastDEM(const FastDEM&) = delete;
  FastDEM& operator=(const FastDEM&) = delete;

  FastDEM& setMappingMode(MappingMode mode) { cfg_.mapping.mode = mode; return push(); }
  FastDEM& setEstimatorType(EstimationType t) { cfg_.mapping.estimation_type = t; return push(); }
  FastDEM& setSensorModel(SensorType t) { cfg_.sensor_model.type = t; return push(); }
  FastDEM& setHeightFilter(float z_min, float z_max) noexcept {
    cfg_.point_filter.z_min = z_min; cfg_.point_filter.z_max = z_max; return push();
  }
  FastDEM& setRangeFilter(float range_min, float range_max) noexcept {
    cfg_.point_filter.range_min = range_min; cfg_.point_filter.range_max = range_max; return push();
  }
  FastDEM& enableRaycasting(bool enabled = true) noexcept { cfg_.raycasting.enabled = enabled; return push(); }
  FastDEM& setCalibrationProvider(std::shared_ptr<Calibration> c) noexcept { calibration_ = std::move(c); return *this; }
  FastDEM& setOdometryProvider(std::shared_ptr<Odometry> o) noexcept { odometry_ = std::move(o); return *this; }
  template <typename T>
  FastDEM& setTransformProvider(std::shared_ptr<T> system) {
    setCalibrationProvider(system);
    setOdometryProvider(system);
    return *this;
  }
  bool hasTransformProvider() const noexcept { return calibration_ != nullptr && odometry_ != nullptr; }
  void reset() { map_.clearAll(); }
  const Config& config() const noexcept { return cfg_; }
  void onScanPreprocessed(CloudCallback cb) { on_preprocessed_ = std::move(cb); }
  void onScanRasterized(CloudCallback cb) { on_rasterized_ = std::move(cb); }

  // fastdem.cpp:83-120
  bool integrate(std::shared_ptr<PointCloud> cloud) {
    if (!calibration_ || !odometry_) {
      std::fprintf(stderr, "[FastDEM] Transform providers not set.\n");
      return false;
    }
    if (!cloud || cloud->empty()) {
      std::fprintf(stderr, "[FastDEM] Received empty or null cloud. Skipping...\n");
      return false;
    }
    if (cloud->frameId().empty()) {
      std::fprintf(stderr, "[FastDEM] Input cloud has no frameId. Skipping...\n");
      return false;
    }
    auto T_base_sensor = calibration_->getExtrinsic(cloud->frameId());
    if (!T_base_sensor) return false;
    auto T_world_base = odometry_->getPoseAt(cloud->timestamp());
    if (!T_world_base) return false;
    return integrateImpl(*cloud, *T_base_sensor, *T_world_base);
  }

  // fastdem.cpp:122-131
  bool integrate(const PointCloud& cloud, const Eigen::Isometry3d& T_base_sensor,
                 const Eigen::Isometry3d& T_world_base) {
    if (cloud.empty()) {
      std::fprintf(stderr, "[FastDEM] Received empty cloud. Skipping...\n");
      return false;
    }
    return integrateImpl(cloud, T_base_sensor, T_world_base);
  }

  const fdem_scan_stats& lastStats() const { return stats_; }

 private:
  FastDEM& push() {
    fdem_config a = toAbi(cfg_);
    check(fdem_mapper_set_config(h_, &a));
    return *this;
  }
  bool integrateImpl(const PointCloud& cloud, const Eigen::Isometry3d& Tbs, const Eigen::Isometry3d& Twb) {
    check(fdem_mapper_integrate(h_, cloud.xyzw(), cloud.intensities(), cloud.colors(), cloud.size(),
                                Tbs.matrix().data(), Twb.matrix().data(), &stats_));
    if (!stats_.integrated) return false;  // all points filtered (fastdem.cpp:138)
    if (on_preprocessed_) {
      int64_t n = 0;
      check(fdem_mapper_last_preprocessed(h_, nullptr, nullptr, nullptr, &n));
      std::vector<float> xyzw(static_cast<size_t>(n > 0 ? n : 1) * 4);
      check(fdem_mapper_last_preprocessed(h_, xyzw.data(), nullptr, nullptr, &n));
      xyzw.resize(static_cast<size_t>(n) * 4);
      for (int64_t i = 0; i < n; ++i) xyzw[4 * i + 3] = 1.0f;  // slot 3 carried sigma_z^2
      PointCloud pc;
      pc.setPointsXYZW(std::move(xyzw));
      pc.setFrameId(map_.getFrameId());
      on_preprocessed_(pc);
    }
    if (on_rasterized_ && stats_.n_cells > 0) {
      int64_t n = 0;
      check(fdem_mapper_last_rasterized(h_, nullptr, &n));
      std::vector<float> xyz(static_cast<size_t>(n > 0 ? n : 1) * 3);
      check(fdem_mapper_last_rasterized(h_, xyz.data(), &n));
      PointCloud pc;
      for (int64_t i = 0; i < n; ++i) pc.add(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      pc.setFrameId(map_.getFrameId());
      on_rasterized_(pc);
    }
    return true;
  }

  ElevationMap& map_;
  Config cfg_;
  fdem_mapper* h_ = nullptr;
  fdem_scan_stats stats_{};
  std::shared_ptr<Calibration> calibration_;
  std::shared_ptr<Odometry> odometry_;
  CloudCallback on_preprocessed_, on_rasterized_;
};

// fastdem::applyRaycasting (fastdem/include/fastdem/postprocess/raycasting.hpp:49-51)
inline void applyRaycasting(ElevationMap& map, const PointCloud& scan, const Eigen::Vector3f& sensor_origin,
                            const config::Raycasting& rc) {
  Config c;
  c.raycasting = rc;
  fdem_config a = toAbi(c);
  const float o[3] = {sensor_origin(0), sensor_origin(1), sensor_origin(2)};
  check(fdem_raycast(map.handle(), scan.xyzw(), scan.size(), o, &a));
}

// fastdem::applyInpainting (fastdem/include/fastdem/postprocess/inpainting.hpp)
inline void applyInpainting(ElevationMap& map, int max_iterations = 3, int min_valid_neighbors = 2,
                            bool inplace = false) {
  check(fdem_inpaint(map.handle(), max_iterations, min_valid_neighbors, inplace ? 1 : 0));
}

// fastdem::applySpatialSmoothing (fastdem/include/fastdem/postprocess/spatial_smoothing.hpp:38-67)
inline void applySpatialSmoothing(ElevationMap& map, const std::string& layer_name, int kernel_size = 3,
                                  int min_valid_neighbors = 5) {
  check(fdem_spatial_smoothing(map.handle(), layer_name.c_str(), kernel_size, min_valid_neighbors));
}

// fastdem::applyUncertaintyFusion (fastdem/src/uncertainty_fusion.cpp:103-186)
inline void applyUncertaintyFusion(ElevationMap& map, const config::UncertaintyFusion& c) {
  if (!c.enabled) return;
  check(fdem_uncertainty_fusion(map.handle(), c.search_radius, c.spatial_sigma, c.quantile_lower,
                                c.quantile_upper, c.min_valid_neighbors));
}

// fastdem::applyFeatureExtraction (fastdem/src/feature_extraction.cpp:28-118)
inline void applyFeatureExtraction(ElevationMap& map, float analysis_radius = 0.3f, int min_valid_neighbors = 4,
                                   float step_lower_percentile = 0.05f, float step_upper_percentile = 0.95f) {
  check(fdem_feature_extraction(map.handle(), analysis_radius, min_valid_neighbors, step_lower_percentile,
                                step_upper_percentile));
}

}  // namespace fastdem
