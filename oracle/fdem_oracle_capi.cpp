// fdem_oracle_capi.cpp — C entry points over the CPU oracle, for ctypes.
// TEST INFRASTRUCTURE ONLY (see fdem_oracle.hpp header).  Loaded exclusively by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>

#include "fdem_oracle.hpp"

using namespace fdem_oracle;

extern "C" {

// Same field order as fdem_config in include/fastdem_b200.h (kept in sync by
// tests/test_abi_layout.py) so one ctypes Structure serves both libraries.
struct orc_config {
  float z_min, z_max, range_min, range_max;
  int32_t sensor_type;
  float lidar_range_noise, lidar_angular_noise;
  float rgbd_normal_a, rgbd_normal_b, rgbd_normal_c, rgbd_lateral_factor;
  float constant_uncertainty;
  int32_t mode;
  int32_t estimation_type;
  float kalman_min_variance, kalman_max_variance, kalman_process_noise;
  float p2_dn[5];
  int32_t p2_elevation_marker;
  float p2_max_sample_count;
  int32_t raycasting_enabled;
  float rc_height_conflict_threshold, rc_log_odds_observed, rc_log_odds_ghost, rc_log_odds_max,
      rc_clear_threshold;
  int32_t move_clear_policy;
};

struct orc_stats {
  int64_t n_input, n_kept, n_cells, n_voxels;
  int32_t integrated;
  int32_t _pad;
};

static Config toConfig(const orc_config* c) {
  Config k;
  if (!c) return k;
  k.z_min = c->z_min;
  k.z_max = c->z_max;
  k.range_min = c->range_min;
  k.range_max = c->range_max;
  k.sensor_type = c->sensor_type;
  k.lidar_range_noise = c->lidar_range_noise;
  k.lidar_angular_noise = c->lidar_angular_noise;
  k.rgbd_normal_a = c->rgbd_normal_a;
  k.rgbd_normal_b = c->rgbd_normal_b;
  k.rgbd_normal_c = c->rgbd_normal_c;
  k.rgbd_lateral_factor = c->rgbd_lateral_factor;
  k.constant_uncertainty = c->constant_uncertainty;
  k.mode = c->mode;
  k.estimation_type = c->estimation_type;
  k.kalman_min_variance = c->kalman_min_variance;
  k.kalman_max_variance = c->kalman_max_variance;
  k.kalman_process_noise = c->kalman_process_noise;
  for (int i = 0; i < 5; ++i) k.p2_dn[i] = c->p2_dn[i];
  k.p2_elevation_marker = c->p2_elevation_marker;
  k.p2_max_sample_count = c->p2_max_sample_count;
  k.raycasting_enabled = c->raycasting_enabled;
  k.rc_height_conflict_threshold = c->rc_height_conflict_threshold;
  k.rc_log_odds_observed = c->rc_log_odds_observed;
  k.rc_log_odds_ghost = c->rc_log_odds_ghost;
  k.rc_log_odds_max = c->rc_log_odds_max;
  k.rc_clear_threshold = c->rc_clear_threshold;
  k.move_clear_policy = c->move_clear_policy;
  return k;
}

void orc_config_default(orc_config* c) {
  Config k;
  c->z_min = k.z_min;
  c->z_max = k.z_max;
  c->range_min = k.range_min;
  c->range_max = k.range_max;
  c->sensor_type = k.sensor_type;
  c->lidar_range_noise = k.lidar_range_noise;
  c->lidar_angular_noise = k.lidar_angular_noise;
  c->rgbd_normal_a = k.rgbd_normal_a;
  c->rgbd_normal_b = k.rgbd_normal_b;
  c->rgbd_normal_c = k.rgbd_normal_c;
  c->rgbd_lateral_factor = k.rgbd_lateral_factor;
  c->constant_uncertainty = k.constant_uncertainty;
  c->mode = k.mode;
  c->estimation_type = k.estimation_type;
  c->kalman_min_variance = k.kalman_min_variance;
  c->kalman_max_variance = k.kalman_max_variance;
  c->kalman_process_noise = k.kalman_process_noise;
  for (int i = 0; i < 5; ++i) c->p2_dn[i] = k.p2_dn[i];
  c->p2_elevation_marker = k.p2_elevation_marker;
  c->p2_max_sample_count = k.p2_max_sample_count;
  c->raycasting_enabled = k.raycasting_enabled;
  c->rc_height_conflict_threshold = k.rc_height_conflict_threshold;
  c->rc_log_odds_observed = k.rc_log_odds_observed;
  c->rc_log_odds_ghost = k.rc_log_odds_ghost;
  c->rc_log_odds_max = k.rc_log_odds_max;
  c->rc_clear_threshold = k.rc_clear_threshold;
  c->move_clear_policy = k.move_clear_policy;
}

static Cloud makeCloud(const float* xyzw, const float* intensity, const uint8_t* rgb, size_t n) {
  Cloud c;
  c.pts.resize(n);
  std::memcpy(static_cast<void*>(c.pts.data()), xyzw, n * 16);
  if (intensity) {
    c.has_intensity = true;
    c.intensity.assign(intensity, intensity + n);
  }
  if (rgb) {
    c.has_color = true;
    c.color.resize(n);
    for (size_t i = 0; i < n; ++i) c.color[i] = Color{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
  }
  return c;
}

static Iso3d toIso(const double* m) {
  Iso3d t;
  std::memcpy(t.m, m, sizeof(t.m));
  return t;
}

// ── map ──────────────────────────────────────────────────────────────────────
void* orc_map_create(float width, float height, float resolution) {
  auto* m = new ElevationMap();
  m->setGeometry(width, height, resolution);
  return m;
}
void orc_map_destroy(void* m) { delete static_cast<ElevationMap*>(m); }
void orc_map_geometry(void* mp, int32_t* rows, int32_t* cols, double* res, double* len2,
                      double* pos2, int32_t* start2) {
  auto* m = static_cast<ElevationMap*>(mp);
  *rows = m->rows();
  *cols = m->cols();
  *res = m->resolution();
  len2[0] = m->length()[0];
  len2[1] = m->length()[1];
  pos2[0] = m->position()[0];
  pos2[1] = m->position()[1];
  start2[0] = m->startIndex().r;
  start2[1] = m->startIndex().c;
}
int orc_map_layer_exists(void* mp, const char* name) {
  return static_cast<ElevationMap*>(mp)->exists(name) ? 1 : 0;
}
int orc_map_layer_count(void* mp) {
  return static_cast<int>(static_cast<ElevationMap*>(mp)->layers().size());
}
int orc_map_layer_name(void* mp, int i, char* buf, int cap) {
  const auto& L = static_cast<ElevationMap*>(mp)->layers();
  if (i < 0 || i >= static_cast<int>(L.size())) return -1;
  std::snprintf(buf, cap, "%s", L[i].c_str());
  return 0;
}
int orc_map_layer_get(void* mp, const char* name, float* dst) {
  auto* m = static_cast<ElevationMap*>(mp);
  if (!m->exists(name)) return -1;
  const Matrix& d = m->get(name);
  std::memcpy(dst, d.data(), d.size() * 4);
  return 0;
}
int orc_map_layer_set(void* mp, const char* name, const float* src) {
  auto* m = static_cast<ElevationMap*>(mp);
  if (!m->exists(name)) m->add(name);
  Matrix& d = m->get(name);
  std::memcpy(d.data(), src, d.size() * 4);
  return 0;
}
int orc_map_layer_add(void* mp, const char* name, float fill) {
  static_cast<ElevationMap*>(mp)->add(name, fill);
  return 0;
}
void orc_map_clear_all(void* mp) { static_cast<ElevationMap*>(mp)->clearAll(); }
int orc_map_is_empty(void* mp) { return static_cast<ElevationMap*>(mp)->isEmpty() ? 1 : 0; }
int orc_map_is_inside(void* mp, double x, double y) {
  return static_cast<ElevationMap*>(mp)->isInside(x, y) ? 1 : 0;
}
int orc_map_get_index(void* mp, double x, double y, int32_t* row, int32_t* col) {
  Index i;
  if (!static_cast<ElevationMap*>(mp)->getIndex(x, y, i)) return 0;
  *row = i.r;
  *col = i.c;
  return 1;
}
void orc_map_get_position(void* mp, int32_t row, int32_t col, double* x, double* y) {
  static_cast<ElevationMap*>(mp)->getPosition(Index{row, col}, *x, *y);
}
int orc_map_move(void* mp, double x, double y, int32_t policy) {
  return static_cast<ElevationMap*>(mp)->move(x, y, policy) ? 1 : 0;
}
void orc_map_set_position(void* mp, double x, double y) {
  static_cast<ElevationMap*>(mp)->setPosition(x, y);
}
void orc_map_set_start_index(void* mp, int32_t r, int32_t c) {
  static_cast<ElevationMap*>(mp)->setStartIndex(Index{r, c});
}

// ── mapper (fastdem::FastDEM) ────────────────────────────────────────────────
struct OrcMapper {
  ElevationMap* map;
  Config cfg;
  std::unique_ptr<FastDEM> dem;
};

void* orc_mapper_create(void* mp, const orc_config* c) {
  auto* h = new OrcMapper();
  h->map = static_cast<ElevationMap*>(mp);
  h->cfg = toConfig(c);
  h->dem = std::make_unique<FastDEM>(*h->map, h->cfg);
  return h;
}
void orc_mapper_destroy(void* h) { delete static_cast<OrcMapper*>(h); }

// FastDEM::integrate(cloud, T_base_sensor, T_world_base).  Returns 1/0 like the
// reference's bool.  `elapsed_s` (optional) times integrate() only — the Cloud is
// built before the clock starts, as the reference receives a ready PointCloud.
int orc_mapper_integrate(void* hp, const float* xyzw, const float* intensity, const uint8_t* rgb,
                         size_t n, const double* T_base_sensor, const double* T_world_base,
                         orc_stats* stats, double* elapsed_s) {
  auto* h = static_cast<OrcMapper*>(hp);
  Cloud cloud = makeCloud(xyzw, intensity, rgb, n);
  ScanStats s;
  const Iso3d Tbs = toIso(T_base_sensor), Twb = toIso(T_world_base);
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = h->dem->integrate(cloud, Tbs, Twb, &s);
  const auto t1 = std::chrono::steady_clock::now();
  if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
  if (stats) {
    stats->n_input = s.n_input;
    stats->n_kept = s.n_kept;
    stats->n_cells = s.n_cells;
    stats->n_voxels = s.n_voxels;
    stats->integrated = s.integrated;
    stats->_pad = 0;
  }
  return ok ? 1 : 0;
}

// ElevationMapping::update(cloud_in_map_frame, robot_xy) — the lower seam that
// tests/test_dual_layer.cpp drives.  `var_z` optional (cloud without covariance
// when null).  Returns the number of touched cells.
int64_t orc_mapper_update(void* hp, const float* xyzw, const float* var_z, const float* intensity,
                          const uint8_t* rgb, size_t n, double robot_x, double robot_y) {
  auto* h = static_cast<OrcMapper*>(hp);
  Cloud cloud = makeCloud(xyzw, intensity, rgb, n);
  if (var_z) {
    cloud.useCovariance();
    for (size_t i = 0; i < n; ++i) cloud.cov[i](2, 2) = var_z[i];
  }
  return static_cast<int64_t>(h->dem->mapping().update(cloud, robot_x, robot_y).size());
}

// ── stage-level entry points (unit / known-answer tests) ─────────────────────

// preprocessScan: returns n_kept; out_xyzw [n*4], out_cov [n*9] (col-major 3x3),
// out_src [n] = original index of each surviving point (stable compaction).
int64_t orc_preprocess(const orc_config* c, const float* xyzw, size_t n, const double* Tbs,
                       const double* Twb, float* out_xyzw, float* out_cov, int32_t* out_src) {
  const Config cfg = toConfig(c);
  Cloud cloud = makeCloud(xyzw, nullptr, nullptr, n);
  // track source indices through the stable compaction via the intensity channel
  cloud.has_intensity = true;
  cloud.intensity.resize(n);
  for (size_t i = 0; i < n; ++i) {
    int32_t v = static_cast<int32_t>(i);
    std::memcpy(&cloud.intensity[i], &v, 4);
  }
  Cloud p = preprocessScan(cfg, cloud, toIso(Tbs), toIso(Twb));
  for (size_t i = 0; i < p.size(); ++i) {
    std::memcpy(out_xyzw + 4 * i, &p.pts[i], 16);
    if (out_cov) std::memcpy(out_cov + 9 * i, p.cov[i].m, 36);
    if (out_src) std::memcpy(&out_src[i], &p.intensity[i], 4);
  }
  return static_cast<int64_t>(p.size());
}

void orc_sensor_cov(const orc_config* c, const float* p, float* out9) {
  const Mat3f m = sensorCov(toConfig(c), p[0], p[1], p[2]);
  std::memcpy(out9, m.m, 36);
}

void orc_transform(const double* T, const float* xyzw, size_t n, float* out) {
  const Mat4f M = castf(toIso(T));
  for (size_t i = 0; i < n; ++i) {
    Vec4f p;
    std::memcpy(&p, xyzw + 4 * i, 16);
    p = mul(M, p);
    std::memcpy(out + 4 * i, &p, 16);
  }
}

// state6 = {x, P, count, sample_mean, sample_var, m2}; out2 = {upper, lower}
void orc_kalman_step(float* state6, float z, float var, float min_v, float max_v, float q,
                     float* out2) {
  Kalman::step(state6[0], state6[1], state6[2], state6[3], state6[4], state6[5], z, var, min_v,
               max_v, q);
  const float sigma = std::sqrt(std::max(0.0f, state6[4]));
  out2[0] = state6[0] + 2.0f * sigma;
  out2[1] = state6[0] - 2.0f * sigma;
}

// P2Quantile::update on a single cell state {q[5], n[5], count}; returns the
// value update() writes to `elevation` (before computeBounds overwrites it).
float orc_p2_step(float* q5, float* n5, float* count, float x, const float* dn5, int marker,
                  float max_count) {
  P2Quantile p(dn5, marker, max_count);
  p.updateP2(q5, n5, *count, x);
  const int m = std::min(std::max(marker, 0), 4);
  return (*count >= 5.0f) ? q5[m] : x;
}

int64_t orc_voxel_any(const float* xyzw, size_t n, float voxel_size, uint32_t* out_idx) {
  Cloud c = makeCloud(xyzw, nullptr, nullptr, n);
  std::vector<uint32_t> sel;
  try {
    sel = voxelGridAnyIndices(c, voxel_size);
  } catch (const std::invalid_argument&) {
    return -1;
  }
  std::memcpy(out_idx, sel.data(), sel.size() * 4);
  return static_cast<int64_t>(sel.size());
}

void orc_raycast(void* mp, const float* xyzw, size_t n, const float* origin3, const orc_config* c) {
  Cloud scan = makeCloud(xyzw, nullptr, nullptr, n);
  applyRaycasting(*static_cast<ElevationMap*>(mp), scan, origin3, toConfig(c));
}

void orc_inpaint(void* mp, int max_iterations, int min_valid_neighbors, int inplace) {
  applyInpainting(*static_cast<ElevationMap*>(mp), max_iterations, min_valid_neighbors,
                  inplace != 0);
}

float orc_pack_color(uint8_t r, uint8_t g, uint8_t b) { return packColor(r, g, b); }

}  // extern "C"

// ── sensor_msgs/PointCloud2 -> PointCloud (nanopcl::from, bridge/ros/impl.hpp:180-270) ──
extern "C" {
struct orc_pc2_layout {
  uint32_t point_step;
  int32_t off_x, off_y, off_z, off_intensity, intensity_type, off_rgb;
};
// Restates from_impl for the channels the path reads: points with a non-finite coordinate are
// skipped; intensity converted per PointField datatype (readIntensity :106-121); rgb unpacked
// from the packed field (readRgb :170-177).  Returns the number of points kept.
int64_t orc_from_pointcloud2(const uint8_t* data, size_t n, const orc_pc2_layout* lo,
                             float* out_xyzw, float* out_intensity, uint8_t* out_rgb) {
  if (n == 0 || lo->off_x < 0 || lo->off_y < 0 || lo->off_z < 0) return 0;
  int64_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint8_t* pt = data + i * lo->point_step;
    float x, y, z;
    std::memcpy(&x, pt + lo->off_x, 4);
    std::memcpy(&y, pt + lo->off_y, 4);
    std::memcpy(&z, pt + lo->off_z, 4);
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    out_xyzw[4 * k + 0] = x;
    out_xyzw[4 * k + 1] = y;
    out_xyzw[4 * k + 2] = z;
    out_xyzw[4 * k + 3] = 1.0f;
    if (lo->off_intensity >= 0 && out_intensity) {
      const uint8_t* p = pt + lo->off_intensity;
      float v = 0.0f;
      switch (lo->intensity_type) {
        case 2: v = static_cast<float>(*p); break;
        case 4: { uint16_t u; std::memcpy(&u, p, 2); v = static_cast<float>(u); break; }
        case 7: std::memcpy(&v, p, 4); break;
        case 8: { double d; std::memcpy(&d, p, 8); v = static_cast<float>(d); break; }
        default: v = 0.0f;
      }
      out_intensity[k] = v;
    }
    if (lo->off_rgb >= 0 && out_rgb) {
      uint32_t rgb;
      std::memcpy(&rgb, pt + lo->off_rgb, 4);
      out_rgb[3 * k + 0] = static_cast<uint8_t>((rgb >> 16) & 0xFF);
      out_rgb[3 * k + 1] = static_cast<uint8_t>((rgb >> 8) & 0xFF);
      out_rgb[3 * k + 2] = static_cast<uint8_t>(rgb & 0xFF);
    }
    ++k;
  }
  return k;
}
}  // extern "C"

extern "C" void orc_spatial_smoothing(void* mp, const char* layer, int kernel_size, int min_valid) {
  applySpatialSmoothing(*static_cast<ElevationMap*>(mp), layer, kernel_size, min_valid);
}

extern "C" void orc_uncertainty_fusion(void* mp, float search_radius, float spatial_sigma,
                                       float quantile_lower, float quantile_upper, int min_valid) {
  UncertaintyFusionConfig c;
  c.enabled = true;
  c.search_radius = search_radius;
  c.spatial_sigma = spatial_sigma;
  c.quantile_lower = quantile_lower;
  c.quantile_upper = quantile_upper;
  c.min_valid_neighbors = min_valid;
  applyUncertaintyFusion(*static_cast<ElevationMap*>(mp), c);
}

extern "C" void orc_feature_extraction(void* mp, float analysis_radius, int min_valid,
                                       float step_lower_percentile, float step_upper_percentile) {
  applyFeatureExtraction(*static_cast<ElevationMap*>(mp), analysis_radius, min_valid,
                         step_lower_percentile, step_upper_percentile);
}

// Eigen-style direct 3x3 symmetric eigen-decomposition (row-major 9 floats in, 3 + 9 out)
extern "C" void orc_eig3(const float* cov9, float* val3, float* vec9) {
  float c[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i][j] = cov9[i * 3 + j];
  const Eig3 e = eig3_direct(c);
  for (int k = 0; k < 3; ++k) {
    val3[k] = e.val[k];
    for (int i = 0; i < 3; ++i) vec9[k * 3 + i] = e.vec[k][i];
  }
}

// map -> PointCloud2: returns the packed size; fills dst (cap bytes) and the field list
// ('\n'-separated names) when they fit
extern "C" int64_t orc_to_pointcloud2(void* mp, const char* elevation_layer, int sub_r0, int sub_c0,
                                      int sub_rows, int sub_cols, uint8_t* dst, int64_t cap,
                                      char* fields, int fields_cap, uint32_t* point_step,
                                      uint32_t* width) {
  const ElevationMap& m = *static_cast<ElevationMap*>(mp);
  const PackedCloud pc = toPointCloud2(m, elevation_layer, Index{sub_r0, sub_c0}, sub_rows, sub_cols);
  *point_step = pc.point_step;
  *width = pc.width;
  std::string names;
  for (size_t i = 0; i < pc.fields.size(); ++i) names += (i ? "\n" : "") + pc.fields[i];
  if (fields && static_cast<int>(names.size()) < fields_cap) std::strcpy(fields, names.c_str());
  if (dst && static_cast<int64_t>(pc.data.size()) <= cap) std::memcpy(dst, pc.data.data(), pc.data.size());
  return static_cast<int64_t>(pc.data.size());
}
