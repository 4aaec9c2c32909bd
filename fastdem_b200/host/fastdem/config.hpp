// config.hpp — fastdem::Config and its parts, field for field
// (fastdem/include/fastdem/config/{fastdem,mapping,sensor_model,postprocess}.hpp).
#pragma once

#include <limits>

#include "fastdem_b200.h"

namespace fastdem {

enum class MappingMode { LOCAL, GLOBAL };
enum class EstimationType { Kalman, P2Quantile };
enum class SensorType { Constant, LiDAR, RGBD };

namespace config {
struct PointFilter {
  float z_min = -std::numeric_limits<float>::max();
  float z_max = std::numeric_limits<float>::max();
  float range_min = 0.0f;
  float range_max = std::numeric_limits<float>::max();
};
struct SensorModel {
  SensorType type = SensorType::LiDAR;
  struct LiDAR { float range_noise = 0.02f; float angular_noise = 0.001f; } lidar;
  struct RGBD { float normal_a = 0.001f; float normal_b = 0.002f; float normal_c = 0.4f; float lateral_factor = 0.001f; } rgbd;
  struct Constant { float uncertainty = 0.03f; } constant;
};
struct Kalman { float min_variance = 0.0001f; float max_variance = 0.01f; float process_noise = 0.0f; };
struct P2Quantile {
  float dn0 = 0.01f, dn1 = 0.16f, dn2 = 0.50f, dn3 = 0.84f, dn4 = 0.99f;
  int elevation_marker = 3;
  float max_sample_count = 0.0f;
};
struct Mapping {
  MappingMode mode = MappingMode::LOCAL;
  EstimationType estimation_type = EstimationType::Kalman;
  Kalman kalman;
  P2Quantile p2;
};
struct Raycasting {
  bool enabled = false;
  float height_conflict_threshold = 0.05f;
  float log_odds_observed = 0.4f;
  float log_odds_ghost = 0.2f;
  float log_odds_max = 2.0f;
  float clear_threshold = -1.0f;
};
// config/postprocess.hpp:25-49
struct Inpainting {
  bool enabled = false;
  int max_iterations = 3;
  int min_valid_neighbors = 2;
};
struct UncertaintyFusion {
  bool enabled = false;
  float search_radius = 0.15f;
  float spatial_sigma = 0.05f;
  float quantile_lower = 0.01f;
  float quantile_upper = 0.99f;
  int min_valid_neighbors = 3;
};
struct FeatureExtraction {
  bool enabled = false;
  float analysis_radius = 0.3f;
  int min_valid_neighbors = 4;
  float step_lower_percentile = 0.05f;
  float step_upper_percentile = 0.95f;
};
}  // namespace config

struct Config {
  config::PointFilter point_filter;
  config::SensorModel sensor_model;
  config::Mapping mapping;
  config::Raycasting raycasting;
};

// flatten into the C-ABI's fdem_config
inline fdem_config toAbi(const Config& c) {
  fdem_config a;
  fdem_config_default(&a);
  a.z_min = c.point_filter.z_min; a.z_max = c.point_filter.z_max;
  a.range_min = c.point_filter.range_min; a.range_max = c.point_filter.range_max;
  a.sensor_type = static_cast<int32_t>(c.sensor_model.type);
  a.lidar_range_noise = c.sensor_model.lidar.range_noise;
  a.lidar_angular_noise = c.sensor_model.lidar.angular_noise;
  a.rgbd_normal_a = c.sensor_model.rgbd.normal_a; a.rgbd_normal_b = c.sensor_model.rgbd.normal_b;
  a.rgbd_normal_c = c.sensor_model.rgbd.normal_c; a.rgbd_lateral_factor = c.sensor_model.rgbd.lateral_factor;
  a.constant_uncertainty = c.sensor_model.constant.uncertainty;
  a.mode = static_cast<int32_t>(c.mapping.mode);
  a.estimation_type = static_cast<int32_t>(c.mapping.estimation_type);
  a.kalman_min_variance = c.mapping.kalman.min_variance;
  a.kalman_max_variance = c.mapping.kalman.max_variance;
  a.kalman_process_noise = c.mapping.kalman.process_noise;
  a.p2_dn[0] = c.mapping.p2.dn0; a.p2_dn[1] = c.mapping.p2.dn1; a.p2_dn[2] = c.mapping.p2.dn2;
  a.p2_dn[3] = c.mapping.p2.dn3; a.p2_dn[4] = c.mapping.p2.dn4;
  a.p2_elevation_marker = c.mapping.p2.elevation_marker;
  a.p2_max_sample_count = c.mapping.p2.max_sample_count;
  a.raycasting_enabled = c.raycasting.enabled ? 1 : 0;
  a.rc_height_conflict_threshold = c.raycasting.height_conflict_threshold;
  a.rc_log_odds_observed = c.raycasting.log_odds_observed;
  a.rc_log_odds_ghost = c.raycasting.log_odds_ghost;
  a.rc_log_odds_max = c.raycasting.log_odds_max;
  a.rc_clear_threshold = c.raycasting.clear_threshold;
  return a;
}

}  // namespace fastdem
