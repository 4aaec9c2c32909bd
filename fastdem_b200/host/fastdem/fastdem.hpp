// fastdem.hpp — fastdem::FastDEM (fastdem/include/fastdem/fastdem.hpp:55-156): same
// constructors, fluent setters, integrate() overloads, callbacks and return/error behaviour
// (fastdem/src/fastdem.cpp), forwarding to libfastdem_b200.so.  The reference's spdlog messages
// become stderr lines.
#pragma once

#include <cstdio>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <string>

#include "fastdem/config.hpp"
#include "fastdem/elevation_map.hpp"
#include "fastdem/point_types.hpp"
#include "fastdem/sensors.hpp"

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
  FastDEM& setSensorModel(SensorType t) {
    cfg_.sensor_model.type = t;
    custom_model_.reset();  // fastdem.cpp:40-44: the setter re-creates the model from the config
    return push();
  }
  // The reference's extension point (fastdem.hpp:80): any SensorModel subclass.  It cannot run on
  // the device, so integrate() evaluates model->computeCovariances() on the host and passes the
  // per-point sensor-frame covariances through fdem_mapper_integrate_with_cov.
  FastDEM& setSensorModel(std::unique_ptr<SensorModel> model) {
    if (model) custom_model_ = std::move(model);
    return *this;
  }
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
  // R * S * R^T in float, tmp = R*S first, 3-term dots as a0 + (a1 + a2): the evaluation order the
  // oracle fixes for fastdem.cpp:184-187
  static Eigen::Matrix3f rotateCov(const float R[9], const Eigen::Matrix3f& S) {
    auto Rm = [&](int r, int c) { return R[c * 3 + r]; };
    float T[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) T[i][j] = Rm(i, 0) * S(0, j) + (Rm(i, 1) * S(1, j) + Rm(i, 2) * S(2, j));
    Eigen::Matrix3f out = Eigen::Matrix3f::Zero();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) out(i, j) = T[i][0] * Rm(j, 0) + (T[i][1] * Rm(j, 1) + T[i][2] * Rm(j, 2));
    return out;
  }

  bool integrateImpl(const PointCloud& cloud, const Eigen::Isometry3d& Tbs, const Eigen::Isometry3d& Twb) {
    if (custom_model_) {
      // preprocessScan's first step on the host (fastdem.cpp:171), the rest on the device
      const PointCloud with_cov = custom_model_->computeCovariances(cloud);
      static_assert(sizeof(Eigen::Matrix3f) == 9 * sizeof(float), "covariances are passed as N x 9 floats");
      check(fdem_mapper_integrate_with_cov(h_, cloud.xyzw(), with_cov.covariances().data()->data(),
                                           cloud.intensities(), cloud.colors(), cloud.size(),
                                           Tbs.matrix().data(), Twb.matrix().data(), &stats_));
    } else {
      check(fdem_mapper_integrate(h_, cloud.xyzw(), cloud.intensities(), cloud.colors(), cloud.size(),
                                  Tbs.matrix().data(), Twb.matrix().data(), &stats_));
    }
    if (!stats_.integrated) return false;  // all points filtered (fastdem.cpp:138)
    if (on_preprocessed_) {
      // the preprocessed cloud in the map frame WITH its covariance channel (fastdem.cpp:139-141,
      // 181-187).  The device keeps only sigma_z^2 = cov(2,2), so the full 3x3 is rebuilt here from
      // the source points: sensor-frame model covariance, rotated by R = (T_wb * T_bs).rotation().
      int64_t n = 0;
      check(fdem_mapper_last_preprocessed(h_, nullptr, nullptr, nullptr, &n));
      std::vector<float> xyzw(static_cast<size_t>(n > 0 ? n : 1) * 4);
      std::vector<int32_t> src(static_cast<size_t>(n > 0 ? n : 1));
      check(fdem_mapper_last_preprocessed(h_, xyzw.data(), nullptr, src.data(), &n));
      xyzw.resize(static_cast<size_t>(n) * 4);
      for (int64_t i = 0; i < n; ++i) xyzw[4 * i + 3] = 1.0f;  // slot 3 carried sigma_z^2
      PointCloud pc;
      pc.setPointsXYZW(std::move(xyzw));
      pc.useCovariance();
      double M[16];
      const double* a = Twb.matrix().data();
      const double* b = Tbs.matrix().data();
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          M[j * 4 + i] = (a[0 * 4 + i] * b[j * 4 + 0] + a[1 * 4 + i] * b[j * 4 + 1]) + a[2 * 4 + i] * b[j * 4 + 2];
      float R[9];
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) R[c * 3 + r] = static_cast<float>(M[c * 4 + r]);
      const std::unique_ptr<SensorModel> builtin = custom_model_ ? nullptr : createSensorModel(cfg_.sensor_model);
      const SensorModel& model = custom_model_ ? *custom_model_ : *builtin;
      for (int64_t i = 0; i < n; ++i)
        pc.covariance(static_cast<size_t>(i)) = rotateCov(R, model.computeCovariance(cloud.point(static_cast<size_t>(src[i]))));
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
  std::unique_ptr<SensorModel> custom_model_;  // null: the built-in model of cfg_.sensor_model runs on the device
  CloudCallback on_preprocessed_, on_rasterized_;
};

// fastdem::ElevationMapping (fastdem/include/fastdem/mapping/elevation_mapping.hpp:20-64): the
// lower seam the reference's test_dual_layer.cpp drives directly — points already in the map
// frame, cloud.covariance(i)(2,2) as the measurement variance, robot position for LOCAL maps.
class ElevationMapping {
 public:
  struct CellObservation {  // elevation_mapping.hpp:26-34; the device reports min_z per touched cell
    float min_z = std::numeric_limits<float>::max();
    float min_z_var = 0.0f;
    float max_z = std::numeric_limits<float>::lowest();
    float max_intensity = std::numeric_limits<float>::lowest();
    float color_packed = 0.0f;
    bool has_intensity = false;
    bool has_color = false;
  };
  // CellMap<CellObservation> of the reference is an unordered_map keyed by Index; iteration order
  // is unspecified there, a vector of (cell centre, observation) here
  struct Entry {
    nanogrid::Position position;
    CellObservation obs;
  };
  using CellObservations = std::vector<Entry>;

  ElevationMapping(ElevationMap& map, const config::Mapping& cfg) : map_(map) {
    Config c;
    c.mapping = cfg;
    fdem_config a = toAbi(c);
    check(fdem_mapper_create(map_.handle(), &a, &h_));  // ensureLayers + obstacle (elevation_mapping.cpp:11-39)
  }
  ~ElevationMapping() { if (h_) fdem_mapper_destroy(h_); }
  ElevationMapping(const ElevationMapping&) = delete;
  ElevationMapping& operator=(const ElevationMapping&) = delete;

  // rasterize + estimate (+ min/max, obstacle, intensity, colour) in one call (elevation_mapping.cpp:110-125)
  CellObservations update(const PointCloud& cloud, const Eigen::Vector2d& robot_position) {
    std::vector<float> var_z;
    if (cloud.hasCovariance()) {
      var_z.resize(cloud.size());
      for (size_t i = 0; i < cloud.size(); ++i) var_z[i] = cloud.covariance(i)(2, 2);  // :58-60
    }
    fdem_scan_stats st{};
    check(fdem_mapper_update(h_, cloud.xyzw(), var_z.empty() ? nullptr : var_z.data(), cloud.intensities(),
                             cloud.colors(), cloud.size(), robot_position(0), robot_position(1), &st));
    CellObservations out;
    if (st.n_cells > 0) {
      int64_t n = 0;
      check(fdem_mapper_last_rasterized(h_, nullptr, &n));
      std::vector<float> xyz(static_cast<size_t>(n > 0 ? n : 1) * 3);
      check(fdem_mapper_last_rasterized(h_, xyz.data(), &n));
      out.resize(static_cast<size_t>(n));
      for (int64_t i = 0; i < n; ++i) {
        out[i].position = nanogrid::Position(xyz[3 * i], xyz[3 * i + 1]);
        out[i].obs.min_z = xyz[3 * i + 2];
        out[i].obs.has_intensity = cloud.hasIntensity();
        out[i].obs.has_color = cloud.hasColor();
      }
    }
    return out;
  }

 private:
  ElevationMap& map_;
  fdem_mapper* h_ = nullptr;
};

inline std::unique_ptr<ElevationMapping> createElevationMapping(ElevationMap& map, const config::Mapping& cfg) {
  return std::make_unique<ElevationMapping>(map, cfg);
}

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
