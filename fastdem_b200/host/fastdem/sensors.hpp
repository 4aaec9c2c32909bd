// sensors.hpp — fastdem::SensorModel and the three built-in models
// (fastdem/include/fastdem/sensors/{sensor_model,lidar_model,rgbd_model}.hpp, src/sensor_model.cpp).
//
// The built-in models run INSIDE the CUDA path (K1 evaluates them per point, selected by enum).
// These host classes exist for the reference's one true extension point,
// FastDEM::setSensorModel(std::unique_ptr<SensorModel>) (fastdem.hpp:80): a user subclass cannot
// run on the device, so the shell evaluates computeCovariances() on the host and hands the
// per-point sensor-frame covariances to fdem_mapper_integrate_with_cov.  The built-ins below use
// the reference's expressions in the oracle's evaluation order, so a built-in model passed as a
// unique_ptr gives the same map as the same model selected by enum.
#pragma once

#include <algorithm>
#include <cmath>
#include <memory>

#include "fastdem/config.hpp"
#include "fastdem/point_types.hpp"

namespace fastdem {

class SensorModel {
 public:
  virtual ~SensorModel() = default;
  // 3x3 covariance in the sensor frame (sensor_model.hpp:46-47)
  virtual Eigen::Matrix3f computeCovariance(const Eigen::Vector3f& point_sensor) const = 0;
  // covariance channel for a whole cloud, taken by value like the reference (:76-85)
  virtual PointCloud computeCovariances(PointCloud scan) const {
    scan.useCovariance();
    auto& covs = scan.covariances();
    for (size_t i = 0; i < scan.size(); ++i) covs[i] = computeCovariance(scan.point(i));
    return scan;
  }
};

class ConstantUncertaintyModel : public SensorModel {  // sensor_model.hpp:62-93
 public:
  explicit ConstantUncertaintyModel(float uncertainty = 0.1f) : variance_(uncertainty * uncertainty) {}
  Eigen::Matrix3f computeCovariance(const Eigen::Vector3f&) const override {
    return Eigen::Matrix3f::Identity() * variance_;
  }

 private:
  float variance_;
};

class LiDARSensorModel : public SensorModel {  // lidar_model.hpp:40-89
 public:
  explicit LiDARSensorModel(float range_noise = 0.02f, float angular_noise = 0.001f)
      : range_noise_(std::fabs(range_noise)), angular_noise_(std::fabs(angular_noise)) {}
  Eigen::Matrix3f computeCovariance(const Eigen::Vector3f& p) const override {
    const float dist_sq = p(0) * p(0) + (p(1) * p(1) + p(2) * p(2));
    if (dist_sq < 1e-6f) return Eigen::Matrix3f::Identity() * 0.01f;
    const float distance = std::sqrt(dist_sq);
    const float dir[3] = {p(0) / distance, p(1) / distance, p(2) / distance};
    const float var_radial = std::max(range_noise_ * range_noise_, 1e-6f);
    const float da = distance * angular_noise_;
    const float var_lateral = std::max(da * da, 1e-6f);
    const float s = var_radial - var_lateral;
    Eigen::Matrix3f cov = Eigen::Matrix3f::Identity() * var_lateral;
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) cov(i, j) = cov(i, j) + dir[j] * (s * dir[i]);
    return cov;
  }

 private:
  float range_noise_, angular_noise_;
};

class RGBDSensorModel : public SensorModel {  // rgbd_model.hpp:50-101
 public:
  explicit RGBDSensorModel(float normal_a = 0.001f, float normal_b = 0.002f, float normal_c = 0.4f,
                           float lateral_factor = 0.001f)
      : a_(normal_a), b_(normal_b), c_(normal_c), k_(lateral_factor) {}
  Eigen::Matrix3f computeCovariance(const Eigen::Vector3f& p) const override {
    const float depth = p(2);
    if (depth <= 0.0f) return Eigen::Matrix3f::Identity() * 0.01f;
    const float diff = depth - c_;
    const float sigma_norm = a_ + b_ * diff * diff;
    const float sigma_lat = k_ * depth;
    Eigen::Matrix3f cov = Eigen::Matrix3f::Zero();
    cov(0, 0) = cov(1, 1) = sigma_lat * sigma_lat;
    cov(2, 2) = sigma_norm * sigma_norm;
    return cov;
  }

 private:
  float a_, b_, c_, k_;
};

// src/sensor_model.cpp:22-40
inline std::unique_ptr<SensorModel> createSensorModel(const config::SensorModel& cfg) {
  switch (cfg.type) {
    case SensorType::LiDAR:
      return std::make_unique<LiDARSensorModel>(cfg.lidar.range_noise, cfg.lidar.angular_noise);
    case SensorType::RGBD:
      return std::make_unique<RGBDSensorModel>(cfg.rgbd.normal_a, cfg.rgbd.normal_b, cfg.rgbd.normal_c,
                                               cfg.rgbd.lateral_factor);
    case SensorType::Constant:
    default:
      return std::make_unique<ConstantUncertaintyModel>(cfg.constant.uncertainty);
  }
}

}  // namespace fastdem
