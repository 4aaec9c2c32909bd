// capi.cu — implementation of include/fastdem_b200.h: device-resident ElevationMap,
// FastDEM mapper, and the per-scan kernel pipeline.  Host logic only; all arithmetic on
// map state happens in kernels.cu / kernels_raycast.cu.  There is no CPU fallback: every
// entry point needs a working CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "../../include/fastdem_b200.h"
#include "device_types.h"

using namespace fdem;

// ───────────────────────────── error plumbing ────────────────────────────────

static thread_local std::string g_last_error;

static fdem_status set_error(fdem_status s, const std::string& msg) {
  g_last_error = msg;
  return s;
}

#define FDEM_CUDA_TRY(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      char _buf[512];                                                                    \
      std::snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr,                  \
                    cudaGetErrorString(_e), __FILE__, __LINE__);                         \
      (void)cudaGetLastError();                                                          \
      return set_error(_e == cudaErrorMemoryAllocation ? FDEM_ERR_OUT_OF_MEMORY          \
                                                       : FDEM_ERR_CUDA,                  \
                       _buf);                                                            \
    }                                                                                    \
  } while (0)

#define FDEM_TRY(expr)                    \
  do {                                    \
    fdem_status _s = (expr);              \
    if (_s != FDEM_OK) return _s;         \
  } while (0)

#define FDEM_REQUIRE(cond, msg) \
  do {                          \
    if (!(cond)) return set_error(FDEM_ERR_INVALID_ARGUMENT, msg); \
  } while (0)

// ───────────────────────────── handles ───────────────────────────────────────

namespace {

constexpr float kNaN = std::numeric_limits<float>::quiet_NaN();

struct Layer {
  std::string name;
  float* d = nullptr;
};

// what the last kernel of a scan leaves for the host (copied D2H into pinned memory)
struct ScanResult {
  uint32_t counters[CNT_COUNT];
  DeviceState state;
};

}  // namespace

struct fdem_map {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  GridGeom geom{};            // host mirror; valid when !geom_stale
  bool geom_stale = false;    // async scans in flight may have moved the window
  DeviceState* d_state = nullptr;  // [2]
  int parity = 0;
  size_t cells = 0;           // rows_local * cols
  std::vector<Layer> layers;
  // touched-cell list of the last observing scan (obstacle reset)
  uint32_t* d_touched_keys = nullptr;
  float* d_touched_minz = nullptr;
  size_t touched_cap = 0;
  bool obstacle_full_clear = false;  // obstacle was written behind the mapper's back
  // raycasting scratch (allocated on first use)
  uint32_t* d_ray_min_enc = nullptr;
  uint32_t* d_hits = nullptr;
  // small scratch
  uint32_t* d_flag = nullptr;
  ScanResult* h_result = nullptr;  // pinned
  LaunchCounter lc;
};

struct fdem_mapper {
  fdem_map* map = nullptr;
  fdem_config cfg{};
  size_t cap = 0;  // scratch capacity in points
  float4* d_in_xyzw = nullptr;
  float* d_in_intensity = nullptr;
  uint8_t* d_in_rgb = nullptr;
  float* d_in_aux = nullptr;  // cov9 (N x 9) or var_z (N)
  float4* d_pm = nullptr;
  uint32_t *d_keys = nullptr, *d_vals = nullptr, *d_skeys = nullptr, *d_svals = nullptr;
  uint64_t *d_vkeys = nullptr, *d_svkeys = nullptr;  // voxel keys (raycasting)
  uint32_t* d_sel = nullptr;                         // voxel representatives
  void* d_sort_temp = nullptr;
  size_t sort_temp_bytes = 0;
  uint32_t* d_counters = nullptr;
  fdem_scan_stats last{};
  uint32_t last_n = 0;
  bool last_had_work = false;
  bool pending = false;  // async scans queued since the last wait
};

namespace {

struct DeviceGuard {
  int prev = 0;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (ok) cudaSetDevice(prev);
  }
};

Layer* find_layer(fdem_map* m, const char* name) {
  for (auto& l : m->layers)
    if (l.name == name) return &l;
  return nullptr;
}

float* layer_ptr(fdem_map* m, const char* name) {
  Layer* l = find_layer(m, name);
  return l ? l->d : nullptr;
}

LayerTable layer_table(fdem_map* m) {
  LayerTable lt{};
  lt.count = static_cast<int32_t>(m->layers.size());
  for (int i = 0; i < lt.count; ++i) {
    lt.ptr[i] = m->layers[i].d;
    if (m->layers[i].name == "elevation") lt.basic[0] = i;
    if (m->layers[i].name == "elevation_min") lt.basic[1] = i;
    if (m->layers[i].name == "elevation_max") lt.basic[2] = i;
  }
  return lt;
}

fdem_status add_layer(fdem_map* m, const char* name, float fill) {
  Layer* l = find_layer(m, name);
  if (!l) {
    if (m->layers.size() >= static_cast<size_t>(kMaxLayers))
      return set_error(FDEM_ERR_UNSUPPORTED, "too many layers");
    Layer nl;
    nl.name = name;
    FDEM_CUDA_TRY(cudaMalloc(&nl.d, std::max<size_t>(m->cells, 1) * sizeof(float)));
    m->layers.push_back(nl);
    l = &m->layers.back();
  }
  launch_fill(l->d, m->cells, fill, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  if (l->name == "obstacle") m->obstacle_full_clear = true;
  return FDEM_OK;
}

fdem_status ensure_layer(fdem_map* m, const char* name, float fill) {
  if (find_layer(m, name)) return FDEM_OK;
  return add_layer(m, name, fill);
}

// bring the host mirror of the geometry up to date with the device
fdem_status refresh_geometry(fdem_map* m) {
  if (!m->geom_stale) return FDEM_OK;
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  DeviceState st;
  FDEM_CUDA_TRY(cudaMemcpy(&st, m->d_state + m->parity, sizeof(st), cudaMemcpyDeviceToHost));
  m->geom = st.geom;
  m->geom_stale = false;
  return FDEM_OK;
}

fdem_status push_state(fdem_map* m, uint32_t touched_count) {
  DeviceState st{};
  st.geom = m->geom;
  st.touched_count = touched_count;
  FDEM_CUDA_TRY(cudaMemcpyAsync(m->d_state + m->parity, &st, sizeof(st), cudaMemcpyHostToDevice,
                                m->stream));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  return FDEM_OK;
}

bool is_device_pointer(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int key_bits(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

// Isometry3d product, linear part and translation (column-major double[16]);
// 3-term dots left to right — the same order the oracle uses.
void compose(const double* a, const double* b, double* r) {
  auto A = [&](int i, int j) { return a[j * 4 + i]; };
  auto B = [&](int i, int j) { return b[j * 4 + i]; };
  std::memset(r, 0, 16 * sizeof(double));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) r[j * 4 + i] = (A(i, 0) * B(0, j) + A(i, 1) * B(1, j)) + A(i, 2) * B(2, j);
    r[12 + i] = ((A(i, 0) * B(0, 3) + A(i, 1) * B(1, 3)) + A(i, 2) * B(2, 3)) + A(i, 3);
  }
  r[15] = 1.0;
}

fdem_status ensure_estimator_layers(fdem_map* m, const fdem_config& cfg) {
  if (cfg.estimation_type == FDEM_EST_P2QUANTILE) {
    // P2Quantile::ensureLayers (mapping/quantile_estimation.hpp:97-115)
    FDEM_TRY(ensure_layer(m, "variance", kNaN));
    FDEM_TRY(ensure_layer(m, "n_points", 0.0f));
    const char* q[5] = {"_p2_q0", "_p2_q1", "_p2_q2", "_p2_q3", "_p2_q4"};
    const char* n[5] = {"_p2_n0", "_p2_n1", "_p2_n2", "_p2_n3", "_p2_n4"};
    for (int i = 0; i < 5; ++i) FDEM_TRY(ensure_layer(m, q[i], kNaN));
    for (int i = 0; i < 5; ++i) FDEM_TRY(ensure_layer(m, n[i], static_cast<float>(i)));
    FDEM_TRY(ensure_layer(m, "upper_bound", kNaN));
    FDEM_TRY(ensure_layer(m, "lower_bound", kNaN));
  } else {
    // Kalman::ensureLayers (mapping/kalman_estimation.hpp:64-82)
    FDEM_TRY(ensure_layer(m, "variance", 0.0f));
    FDEM_TRY(ensure_layer(m, "n_points", 0.0f));
    FDEM_TRY(ensure_layer(m, "_kalman_p", 0.0f));
    FDEM_TRY(ensure_layer(m, "_sample_mean", kNaN));
    FDEM_TRY(ensure_layer(m, "_sample_m2", 0.0f));
    FDEM_TRY(ensure_layer(m, "upper_bound", kNaN));
    FDEM_TRY(ensure_layer(m, "lower_bound", kNaN));
  }
  // ElevationMapping ctor (elevation_mapping.cpp:38)
  FDEM_TRY(ensure_layer(m, "obstacle", kNaN));
  return FDEM_OK;
}

fdem_status validate_config(const fdem_config* c) {
  FDEM_REQUIRE(c != nullptr, "config is null");
  FDEM_REQUIRE(c->sensor_type >= 0 && c->sensor_type <= 2, "bad sensor_type");
  FDEM_REQUIRE(c->mode == FDEM_MODE_LOCAL || c->mode == FDEM_MODE_GLOBAL, "bad mode");
  FDEM_REQUIRE(c->estimation_type == FDEM_EST_KALMAN || c->estimation_type == FDEM_EST_P2QUANTILE,
               "bad estimation_type");
  FDEM_REQUIRE(c->move_clear_policy == FDEM_MOVE_CLEAR_ALL_LAYERS ||
                   c->move_clear_policy == FDEM_MOVE_CLEAR_BASIC_LAYERS,
               "bad move_clear_policy");
  return FDEM_OK;
}

void free_scratch(fdem_mapper* mp) {
  cudaFree(mp->d_in_xyzw);
  cudaFree(mp->d_in_intensity);
  cudaFree(mp->d_in_rgb);
  cudaFree(mp->d_in_aux);
  cudaFree(mp->d_pm);
  cudaFree(mp->d_keys);
  cudaFree(mp->d_vals);
  cudaFree(mp->d_skeys);
  cudaFree(mp->d_svals);
  cudaFree(mp->d_vkeys);
  cudaFree(mp->d_svkeys);
  cudaFree(mp->d_sel);
  cudaFree(mp->d_sort_temp);
  mp->d_in_xyzw = nullptr; mp->d_in_intensity = nullptr; mp->d_in_rgb = nullptr;
  mp->d_in_aux = nullptr; mp->d_pm = nullptr; mp->d_keys = mp->d_vals = nullptr;
  mp->d_skeys = mp->d_svals = nullptr; mp->d_vkeys = mp->d_svkeys = nullptr;
  mp->d_sel = nullptr; mp->d_sort_temp = nullptr;
  mp->cap = 0;
  mp->sort_temp_bytes = 0;
}

fdem_status ensure_capacity(fdem_mapper* mp, size_t n) {
  fdem_map* m = mp->map;
  const bool need_vox = mp->cfg.raycasting_enabled != 0;
  if (n <= mp->cap && (!need_vox || mp->d_vkeys)) {
    return FDEM_OK;
  }
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));  // scratch may still be in use
  const size_t cap = std::max<size_t>(std::max(n, mp->cap), 1024);
  free_scratch(mp);
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_xyzw, cap * sizeof(float4)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_intensity, cap * sizeof(float)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_rgb, cap * 3 + 16));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_aux, cap * 9 * sizeof(float)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_pm, cap * sizeof(float4)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_keys, cap * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_vals, cap * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_skeys, cap * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_svals, cap * sizeof(uint32_t)));
  size_t temp = sort_pairs_u32_temp_bytes(static_cast<uint32_t>(cap), 32);
  if (need_vox) {
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_vkeys, cap * sizeof(uint64_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_svkeys, cap * sizeof(uint64_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_sel, cap * sizeof(uint32_t)));
    temp = std::max(temp, sort_pairs_u64_temp_bytes(static_cast<uint32_t>(cap), 63));
  }
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_sort_temp, temp));
  mp->sort_temp_bytes = temp;
  mp->cap = cap;
  // the touched list lives in the map and must hold one entry per sorted element
  if (m->touched_cap < cap) {
    uint32_t* nk = nullptr;
    float* nz = nullptr;
    FDEM_CUDA_TRY(cudaMalloc(&nk, cap * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMalloc(&nz, cap * sizeof(float)));
    if (m->d_touched_keys && m->touched_cap) {
      FDEM_CUDA_TRY(cudaMemcpy(nk, m->d_touched_keys, m->touched_cap * sizeof(uint32_t),
                               cudaMemcpyDeviceToDevice));
      FDEM_CUDA_TRY(cudaMemcpy(nz, m->d_touched_minz, m->touched_cap * sizeof(float),
                               cudaMemcpyDeviceToDevice));
    }
    cudaFree(m->d_touched_keys);
    cudaFree(m->d_touched_minz);
    m->d_touched_keys = nk;
    m->d_touched_minz = nz;
    m->touched_cap = cap;
  }
  return FDEM_OK;
}

// copy an input channel to the device scratch unless it already lives there
template <typename T>
fdem_status stage(const T* src, T* scratch, size_t count, cudaStream_t s, const T** out) {
  if (!src) {
    *out = nullptr;
    return FDEM_OK;
  }
  if (is_device_pointer(src)) {
    *out = src;
    return FDEM_OK;
  }
  FDEM_CUDA_TRY(cudaMemcpyAsync(scratch, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
  *out = scratch;
  return FDEM_OK;
}

EstLayers est_layers(fdem_map* m) {
  EstLayers L{};
  L.elevation = layer_ptr(m, "elevation");
  L.elevation_min = layer_ptr(m, "elevation_min");
  L.elevation_max = layer_ptr(m, "elevation_max");
  L.variance = layer_ptr(m, "variance");
  L.n_points = layer_ptr(m, "n_points");
  L.upper_bound = layer_ptr(m, "upper_bound");
  L.lower_bound = layer_ptr(m, "lower_bound");
  L.obstacle = layer_ptr(m, "obstacle");
  L.intensity = layer_ptr(m, "intensity");
  L.color = layer_ptr(m, "color");
  L.kalman_p = layer_ptr(m, "_kalman_p");
  L.sample_mean = layer_ptr(m, "_sample_mean");
  L.sample_m2 = layer_ptr(m, "_sample_m2");
  const char* q[5] = {"_p2_q0", "_p2_q1", "_p2_q2", "_p2_q3", "_p2_q4"};
  const char* n[5] = {"_p2_n0", "_p2_n1", "_p2_n2", "_p2_n3", "_p2_n4"};
  for (int i = 0; i < 5; ++i) {
    L.p2_q[i] = layer_ptr(m, q[i]);
    L.p2_n[i] = layer_ptr(m, n[i]);
  }
  return L;
}

fdem_status raycast_device(fdem_map* m, const fdem_config& cfg, const float origin[3],
                           const float4* pts, const uint32_t* sel, const uint32_t* n_sel_dev,
                           uint32_t n_max, uint32_t* counters);

// ── one scan: K1 -> K2 -> sort -> K3 (-> voxel + raycast) -> result D2H ─────────
struct ScanInputs {
  const float* xyzw;
  const float* intensity;
  const uint8_t* rgb;
  const float* cov9;
  const float* var_z;
  size_t n;
  int input_frame;
  const double* Tbs;
  const double* Twb;
  double robot_x, robot_y;
};

fdem_status enqueue_scan(fdem_mapper* mp, const ScanInputs& in) {
  fdem_map* m = mp->map;
  const fdem_config& cfg = mp->cfg;
  cudaStream_t s = m->stream;
  FDEM_REQUIRE(in.n <= 0x7fffffffu, "too many points");
  const uint32_t n = static_cast<uint32_t>(in.n);
  FDEM_REQUIRE(m->cells < 0xffffffffull, "map too large for 32-bit cell keys");
  if (cfg.mode == FDEM_MODE_LOCAL)
    FDEM_REQUIRE(m->geom.row_begin == 0 && m->geom.row_end == m->geom.rows,
                 "LOCAL mapping is not valid on a row stripe");

  FDEM_TRY(ensure_capacity(mp, n));
  // layers the cloud's channels need (updateIntensity / updateColor add them lazily,
  // elevation_mapping.cpp:155,169)
  if (in.intensity) FDEM_TRY(ensure_layer(m, "intensity", kNaN));
  if (in.rgb) FDEM_TRY(ensure_layer(m, "color", kNaN));

  PreprocessParams pp{};
  const float* xyzw_d = nullptr;
  FDEM_TRY(stage(in.xyzw, reinterpret_cast<float*>(mp->d_in_xyzw), static_cast<size_t>(n) * 4, s,
                 &xyzw_d));
  FDEM_REQUIRE((reinterpret_cast<uintptr_t>(xyzw_d) & 15) == 0, "xyzw must be 16-byte aligned");
  pp.xyzw = reinterpret_cast<const float4*>(xyzw_d);
  const float* inten_d = nullptr;
  FDEM_TRY(stage(in.intensity, mp->d_in_intensity, n, s, &inten_d));
  const uint8_t* rgb_d = nullptr;
  FDEM_TRY(stage(in.rgb, mp->d_in_rgb, static_cast<size_t>(n) * 3, s, &rgb_d));
  FDEM_REQUIRE(!(in.cov9 && in.var_z), "cov9 and var_z are exclusive");
  const float* aux_d = nullptr;
  if (in.cov9) FDEM_TRY(stage(in.cov9, mp->d_in_aux, static_cast<size_t>(n) * 9, s, &aux_d));
  if (in.var_z) FDEM_TRY(stage(in.var_z, mp->d_in_aux, n, s, &aux_d));
  pp.intensity = inten_d;
  pp.cov9 = in.cov9 ? aux_d : nullptr;
  pp.var_z = in.var_z ? aux_d : nullptr;
  pp.n = n;
  pp.input_frame = in.input_frame;
  double T[16];
  if (in.input_frame == INPUT_SENSOR_FRAME) {
    for (int i = 0; i < 16; ++i) pp.T1[i] = static_cast<float>(in.Tbs[i]);
    for (int i = 0; i < 16; ++i) pp.T2[i] = static_cast<float>(in.Twb[i]);
    compose(in.Twb, in.Tbs, T);
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) pp.R[c * 3 + r] = static_cast<float>(T[c * 4 + r]);
  }
  pp.robot_x = in.robot_x;
  pp.robot_y = in.robot_y;
  pp.z_min = cfg.z_min;
  pp.z_max = cfg.z_max;
  pp.range_min_sq = cfg.range_min * cfg.range_min;  // crop_impl.hpp:83-84
  pp.range_max_sq = cfg.range_max * cfg.range_max;  // FLT_MAX^2 = +inf
  pp.sensor_type = cfg.sensor_type;
  pp.lidar_range_noise = std::fabs(cfg.lidar_range_noise);      // lidar_model.hpp:59-62
  pp.lidar_angular_noise = std::fabs(cfg.lidar_angular_noise);
  pp.rgbd_a = cfg.rgbd_normal_a;
  pp.rgbd_b = cfg.rgbd_normal_b;
  pp.rgbd_c = cfg.rgbd_normal_c;
  pp.rgbd_k = cfg.rgbd_lateral_factor;
  pp.constant_variance = cfg.constant_uncertainty * cfg.constant_uncertainty;
  pp.local_mode = cfg.mode == FDEM_MODE_LOCAL ? 1 : 0;
  pp.invalid_key = static_cast<uint32_t>(m->cells);

  const DeviceState* st_in = m->d_state + m->parity;
  DeviceState* st_out = m->d_state + (m->parity ^ 1);

  FDEM_CUDA_TRY(cudaMemsetAsync(mp->d_counters, 0, CNT_COUNT * sizeof(uint32_t), s));
  launch_preprocess_bin(pp, st_in, mp->d_counters, mp->d_pm, mp->d_keys, mp->d_vals, s, m->lc);

  if (m->obstacle_full_clear) {
    // obstacle was uploaded / edited by the caller: fall back to the reference's whole-layer
    // clear (elevation_mapping.cpp:146) until a scan with observations has gone through
    float* ob = layer_ptr(m, "obstacle");
    if (ob) {
      // done unconditionally: if this scan ends up without observations the reference would
      // not clear, but then the layer content was caller-provided and undefined for the path
      launch_fill(ob, m->cells, kNaN, s, m->lc);
    }
  }
  CommitParams cp{};
  cp.robot_x = in.robot_x;
  cp.robot_y = in.robot_y;
  cp.local_mode = pp.local_mode;
  cp.clear_policy = cfg.move_clear_policy;
  cp.invalid_key = pp.invalid_key;
  cp.obstacle = layer_ptr(m, "obstacle");
  cp.touched_keys = m->d_touched_keys;
  launch_commit(cp, st_in, st_out, mp->d_counters, layer_table(m), s, m->lc);

  const int bits = key_bits(m->cells);  // keys are in [0, cells]; `cells` = dropped point
  FDEM_CUDA_TRY(sort_pairs_u32(mp->d_sort_temp, mp->sort_temp_bytes, mp->d_keys, mp->d_skeys,
                               mp->d_vals, mp->d_svals, n, bits, s, m->lc));

  EstimateParams ep{};
  ep.sorted_keys = mp->d_skeys;
  ep.sorted_vals = mp->d_svals;
  ep.pm = mp->d_pm;
  ep.intensity = inten_d;
  ep.rgb = rgb_d;
  ep.touched_keys = m->d_touched_keys;
  ep.touched_minz = m->d_touched_minz;
  ep.n_sorted = n;
  ep.invalid_key = pp.invalid_key;
  ep.estimation_type = cfg.estimation_type;
  ep.kalman_min_variance = cfg.kalman_min_variance;
  ep.kalman_max_variance = cfg.kalman_max_variance;
  ep.kalman_process_noise = cfg.kalman_process_noise;
  {
    // P2Quantile ctor (quantile_estimation.hpp:83-94): clamp to [0,1], enforce monotone
    float dn[5];
    for (int i = 0; i < 5; ++i) dn[i] = std::min(std::max(cfg.p2_dn[i], 0.0f), 1.0f);
    for (int i = 1; i < 5; ++i) dn[i] = std::max(dn[i], dn[i - 1]);
    for (int i = 0; i < 5; ++i) ep.p2_dn[i] = dn[i];
    ep.p2_marker = std::min(std::max(cfg.p2_elevation_marker, 0), 4);
    ep.p2_max_sample_count = std::max(cfg.p2_max_sample_count, 0.0f);
  }
  ep.L = est_layers(m);
  launch_segreduce_estimate(ep, mp->d_counters, mp->d_counters, s, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());

  m->parity ^= 1;
  m->geom_stale = true;

  if (cfg.raycasting_enabled && in.input_frame == INPUT_SENSOR_FRAME) {
    // fastdem.cpp:153-159: sensor origin = (T_world_base*T_base_sensor).translation(),
    // ray_scan = voxelGrid(points, resolution, ANY)
    const float origin[3] = {static_cast<float>(T[12]), static_cast<float>(T[13]),
                             static_cast<float>(T[14])};
    const float voxel = static_cast<float>(m->geom.res);
    if (voxel < 0.001f || voxel > 100.0f)
      return set_error(FDEM_ERR_INVALID_ARGUMENT, "voxel_size must be in [0.001, 100]");
    launch_voxel_keys(mp->d_pm, n, 1.0f / voxel, mp->d_vkeys, mp->d_vals, s, m->lc);
    FDEM_CUDA_TRY(sort_pairs_u64(mp->d_sort_temp, mp->sort_temp_bytes, mp->d_vkeys, mp->d_svkeys,
                                 mp->d_vals, mp->d_svals, n, 64, s, m->lc));
    launch_voxel_select(mp->d_svkeys, mp->d_svals, n, mp->d_counters, mp->d_sel, s, m->lc);
    FDEM_TRY(raycast_device(m, cfg, origin, mp->d_pm, mp->d_sel, mp->d_counters + CNT_VOXELS, n,
                            mp->d_counters));
  }

  // leave the scan's counters + committed state where the host can pick them up
  FDEM_CUDA_TRY(cudaMemcpyAsync(m->h_result->counters, mp->d_counters,
                                CNT_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  FDEM_CUDA_TRY(cudaMemcpyAsync(&m->h_result->state, m->d_state + m->parity, sizeof(DeviceState),
                                cudaMemcpyDeviceToHost, s));
  mp->last_n = n;
  mp->last_had_work = true;
  mp->pending = true;
  return FDEM_OK;
}

fdem_status finish_scan(fdem_mapper* mp, fdem_scan_stats* stats) {
  fdem_map* m = mp->map;
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (mp->last_had_work) {
    const ScanResult& r = *m->h_result;
    m->geom = r.state.geom;
    m->geom_stale = false;
    mp->last.n_input = mp->last_n;
    mp->last.n_kept = r.counters[CNT_KEPT];
    mp->last.n_cells = r.counters[CNT_CELLS];
    mp->last.n_voxels = r.counters[CNT_VOXELS];
    mp->last.integrated = r.counters[CNT_KEPT] > 0 ? 1 : 0;
    mp->last._pad = 0;
    if (mp->last.n_cells > 0) m->obstacle_full_clear = false;
  }
  mp->pending = false;
  if (stats) *stats = mp->last;
  return FDEM_OK;
}

}  // namespace

// ───────────────────────────── library ───────────────────────────────────────

extern "C" {

int32_t fdem_abi_version(void) { return FDEM_ABI_VERSION; }

const char* fdem_last_error(void) { return g_last_error.c_str(); }

const char* fdem_status_string(fdem_status s) {
  switch (s) {
    case FDEM_OK: return "ok";
    case FDEM_ERR_INVALID_ARGUMENT: return "invalid argument";
    case FDEM_ERR_CUDA: return "CUDA error";
    case FDEM_ERR_NO_LAYER: return "no such layer";
    case FDEM_ERR_OUT_OF_MEMORY: return "out of memory";
    case FDEM_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown";
  }
}

void fdem_config_default(fdem_config* c) {
  if (!c) return;
  std::memset(c, 0, sizeof(*c));
  c->z_min = -std::numeric_limits<float>::max();
  c->z_max = std::numeric_limits<float>::max();
  c->range_min = 0.0f;
  c->range_max = std::numeric_limits<float>::max();
  c->sensor_type = FDEM_SENSOR_LIDAR;
  c->lidar_range_noise = 0.02f;
  c->lidar_angular_noise = 0.001f;
  c->rgbd_normal_a = 0.001f;
  c->rgbd_normal_b = 0.002f;
  c->rgbd_normal_c = 0.4f;
  c->rgbd_lateral_factor = 0.001f;
  c->constant_uncertainty = 0.03f;
  c->mode = FDEM_MODE_LOCAL;
  c->estimation_type = FDEM_EST_KALMAN;
  c->kalman_min_variance = 0.0001f;
  c->kalman_max_variance = 0.01f;
  c->kalman_process_noise = 0.0f;
  const float dn[5] = {0.01f, 0.16f, 0.50f, 0.84f, 0.99f};
  for (int i = 0; i < 5; ++i) c->p2_dn[i] = dn[i];
  c->p2_elevation_marker = 3;
  c->p2_max_sample_count = 0.0f;
  c->raycasting_enabled = 0;
  c->rc_height_conflict_threshold = 0.05f;
  c->rc_log_odds_observed = 0.4f;
  c->rc_log_odds_ghost = 0.2f;
  c->rc_log_odds_max = 2.0f;
  c->rc_clear_threshold = -1.0f;
  c->move_clear_policy = FDEM_MOVE_CLEAR_ALL_LAYERS;
}

// ───────────────────────────── map ───────────────────────────────────────────

fdem_status fdem_map_create_stripe(float width, float height, float resolution, int32_t row_begin,
                                   int32_t row_end, int32_t device, void* stream, fdem_map** out) {
  FDEM_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  FDEM_REQUIRE(resolution > 0.0f && width > 0.0f && height > 0.0f, "bad geometry");
  int ndev = 0;
  FDEM_CUDA_TRY(cudaGetDeviceCount(&ndev));
  FDEM_REQUIRE(device >= 0 && device < ndev, "bad device ordinal");
  FDEM_CUDA_TRY(cudaSetDevice(device));

  fdem_map* m = new (std::nothrow) fdem_map();
  if (!m) return set_error(FDEM_ERR_OUT_OF_MEMORY, "host allocation failed");
  m->device = device;
  // ElevationMap::setGeometry(float, float, float) widens to double, then nanoGrid
  // setGeometry: size = round(length / res), length = size * res (elevation_map.hpp:112-116)
  const double res = static_cast<double>(resolution);
  const double L[2] = {static_cast<double>(width), static_cast<double>(height)};
  GridGeom g{};
  g.rows = static_cast<int32_t>(std::round(L[0] / res));
  g.cols = static_cast<int32_t>(std::round(L[1] / res));
  g.res = res;
  g.len[0] = g.rows * res;
  g.len[1] = g.cols * res;
  g.pos[0] = g.pos[1] = 0.0;
  g.start[0] = g.start[1] = 0;
  if (row_begin < 0 && row_end < 0) {
    row_begin = 0;
    row_end = g.rows;
  }
  if (g.rows <= 0 || g.cols <= 0 || row_begin < 0 || row_end > g.rows || row_begin >= row_end) {
    delete m;
    return set_error(FDEM_ERR_INVALID_ARGUMENT, "bad size or row stripe");
  }
  g.row_begin = row_begin;
  g.row_end = row_end;
  m->geom = g;
  m->cells = static_cast<size_t>(row_end - row_begin) * g.cols;

  if (stream) {
    m->stream = static_cast<cudaStream_t>(stream);
  } else {
    FDEM_CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    m->own_stream = true;
  }
  FDEM_CUDA_TRY(cudaMalloc(&m->d_state, 2 * sizeof(DeviceState)));
  FDEM_CUDA_TRY(cudaMemset(m->d_state, 0, 2 * sizeof(DeviceState)));
  FDEM_CUDA_TRY(cudaMalloc(&m->d_flag, 4 * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaHostAlloc(&m->h_result, sizeof(ScanResult), cudaHostAllocDefault));
  std::memset(m->h_result, 0, sizeof(ScanResult));
  fdem_status st = push_state(m, 0);
  if (st != FDEM_OK) return st;
  // ElevationMap() basic layers (elevation_map.hpp:99-103), then clearAll()
  for (const char* name : {"elevation", "elevation_min", "elevation_max"}) {
    st = add_layer(m, name, kNaN);
    if (st != FDEM_OK) return st;
  }
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  *out = m;
  return FDEM_OK;
}

fdem_status fdem_map_create(float width, float height, float resolution, int32_t device,
                            void* stream, fdem_map** out) {
  return fdem_map_create_stripe(width, height, resolution, -1, -1, device, stream, out);
}

fdem_status fdem_map_destroy(fdem_map* m) {
  if (!m) return FDEM_OK;
  DeviceGuard dg(m->device);
  cudaStreamSynchronize(m->stream);
  for (auto& l : m->layers) cudaFree(l.d);
  cudaFree(m->d_state);
  cudaFree(m->d_flag);
  cudaFree(m->d_touched_keys);
  cudaFree(m->d_touched_minz);
  cudaFree(m->d_ray_min_enc);
  cudaFree(m->d_hits);
  cudaFreeHost(m->h_result);
  if (m->own_stream) cudaStreamDestroy(m->stream);
  delete m;
  return FDEM_OK;
}

fdem_status fdem_map_get_geometry(fdem_map* m, fdem_geometry* out) {
  FDEM_REQUIRE(m && out, "null argument");
  DeviceGuard dg(m->device);
  FDEM_TRY(refresh_geometry(m));
  out->rows = m->geom.rows;
  out->cols = m->geom.cols;
  out->resolution = m->geom.res;
  out->length[0] = m->geom.len[0];
  out->length[1] = m->geom.len[1];
  out->position[0] = m->geom.pos[0];
  out->position[1] = m->geom.pos[1];
  out->start_index[0] = m->geom.start[0];
  out->start_index[1] = m->geom.start[1];
  out->row_begin = m->geom.row_begin;
  out->row_end = m->geom.row_end;
  return FDEM_OK;
}

static fdem_status current_touched_count(fdem_map* m, uint32_t* out) {
  DeviceState st;
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  FDEM_CUDA_TRY(cudaMemcpy(&st, m->d_state + m->parity, sizeof(st), cudaMemcpyDeviceToHost));
  *out = st.touched_count;
  return FDEM_OK;
}

fdem_status fdem_map_set_position(fdem_map* m, double x, double y) {
  FDEM_REQUIRE(m, "null map");
  DeviceGuard dg(m->device);
  FDEM_TRY(refresh_geometry(m));
  uint32_t tc = 0;
  FDEM_TRY(current_touched_count(m, &tc));
  m->geom.pos[0] = x;
  m->geom.pos[1] = y;
  return push_state(m, tc);
}

fdem_status fdem_map_set_start_index(fdem_map* m, int32_t row, int32_t col) {
  FDEM_REQUIRE(m, "null map");
  DeviceGuard dg(m->device);
  FDEM_TRY(refresh_geometry(m));
  FDEM_REQUIRE(row >= 0 && row < m->geom.rows && col >= 0 && col < m->geom.cols,
               "start index out of range");
  FDEM_REQUIRE(m->geom.row_begin == 0 && m->geom.row_end == m->geom.rows,
               "start index is fixed at 0 on a row stripe");
  uint32_t tc = 0;
  FDEM_TRY(current_touched_count(m, &tc));
  m->geom.start[0] = row;
  m->geom.start[1] = col;
  return push_state(m, tc);
}

fdem_status fdem_map_move(fdem_map* m, double x, double y, int32_t clear_policy, int32_t* moved) {
  FDEM_REQUIRE(m, "null map");
  FDEM_REQUIRE(clear_policy == 0 || clear_policy == 1, "bad clear policy");
  DeviceGuard dg(m->device);
  FDEM_REQUIRE(m->geom.row_begin == 0 && m->geom.row_end == m->geom.rows,
               "move() is not valid on a row stripe");
  launch_move_only(m->d_state + m->parity, m->d_state + (m->parity ^ 1), x, y, clear_policy,
                   layer_table(m), m->d_flag, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  m->parity ^= 1;
  m->geom_stale = true;
  FDEM_TRY(refresh_geometry(m));
  uint32_t flag = 0;
  FDEM_CUDA_TRY(cudaMemcpy(&flag, m->d_flag, sizeof(flag), cudaMemcpyDeviceToHost));
  if (moved) *moved = flag ? 1 : 0;
  return FDEM_OK;
}

fdem_status fdem_map_is_inside(fdem_map* m, double x, double y, int32_t* inside) {
  FDEM_REQUIRE(m && inside, "null argument");
  DeviceGuard dg(m->device);
  FDEM_TRY(refresh_geometry(m));
  *inside = geom_is_inside(m->geom, x, y) ? 1 : 0;
  return FDEM_OK;
}

fdem_status fdem_map_get_index(fdem_map* m, double x, double y, int32_t* row, int32_t* col,
                               int32_t* inside) {
  FDEM_REQUIRE(m && row && col && inside, "null argument");
  DeviceGuard dg(m->device);
  FDEM_TRY(refresh_geometry(m));
  int32_t r = 0, c = 0;
  *inside = geom_get_index(m->geom, x, y, r, c) ? 1 : 0;
  *row = r;
  *col = c;
  return FDEM_OK;
}

fdem_status fdem_map_get_cell_position(fdem_map* m, int32_t row, int32_t col, double* x,
                                       double* y) {
  FDEM_REQUIRE(m && x && y, "null argument");
  DeviceGuard dg(m->device);
  FDEM_TRY(refresh_geometry(m));
  FDEM_REQUIRE(row >= 0 && row < m->geom.rows && col >= 0 && col < m->geom.cols,
               "index out of range");
  geom_cell_position(m->geom, row, col, *x, *y);
  return FDEM_OK;
}

fdem_status fdem_map_layer_exists(fdem_map* m, const char* name, int32_t* exists) {
  FDEM_REQUIRE(m && name && exists, "null argument");
  *exists = find_layer(m, name) ? 1 : 0;
  return FDEM_OK;
}

fdem_status fdem_map_layer_add(fdem_map* m, const char* name, float fill) {
  FDEM_REQUIRE(m && name && name[0], "null argument");
  DeviceGuard dg(m->device);
  return add_layer(m, name, fill);
}

fdem_status fdem_map_layer_count(fdem_map* m, int32_t* count) {
  FDEM_REQUIRE(m && count, "null argument");
  *count = static_cast<int32_t>(m->layers.size());
  return FDEM_OK;
}

fdem_status fdem_map_layer_name(fdem_map* m, int32_t i, char* buf, int32_t cap) {
  FDEM_REQUIRE(m && buf && cap > 0, "null argument");
  FDEM_REQUIRE(i >= 0 && i < static_cast<int32_t>(m->layers.size()), "layer index out of range");
  std::snprintf(buf, cap, "%s", m->layers[i].name.c_str());
  return FDEM_OK;
}

fdem_status fdem_map_layer_download(fdem_map* m, const char* name, float* dst) {
  FDEM_REQUIRE(m && name && dst, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  FDEM_CUDA_TRY(cudaMemcpyAsync(dst, l->d, m->cells * sizeof(float), cudaMemcpyDefault, m->stream));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  return FDEM_OK;
}

fdem_status fdem_map_layer_upload(fdem_map* m, const char* name, const float* src) {
  FDEM_REQUIRE(m && name && src, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  FDEM_CUDA_TRY(cudaMemcpyAsync(l->d, src, m->cells * sizeof(float), cudaMemcpyDefault, m->stream));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (l->name == "obstacle") m->obstacle_full_clear = true;
  return FDEM_OK;
}

fdem_status fdem_map_layer_device_ptr(fdem_map* m, const char* name, float** dptr) {
  FDEM_REQUIRE(m && name && dptr, "null argument");
  Layer* l = find_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  *dptr = l->d;
  if (l->name == "obstacle") m->obstacle_full_clear = true;  // caller may write through it
  return FDEM_OK;
}

static fdem_status cell_offset(fdem_map* m, int32_t row, int32_t col, int64_t* lin) {
  FDEM_REQUIRE(row >= 0 && row < m->geom.rows && col >= 0 && col < m->geom.cols,
               "index out of range");
  *lin = geom_linear(m->geom, row, col);
  FDEM_REQUIRE(*lin >= 0, "row is outside this handle's stripe");
  return FDEM_OK;
}

fdem_status fdem_map_cell_get(fdem_map* m, const char* name, int32_t row, int32_t col, float* v) {
  FDEM_REQUIRE(m && name && v, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  int64_t lin;
  FDEM_TRY(cell_offset(m, row, col, &lin));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  FDEM_CUDA_TRY(cudaMemcpy(v, l->d + lin, sizeof(float), cudaMemcpyDeviceToHost));
  return FDEM_OK;
}

fdem_status fdem_map_cell_set(fdem_map* m, const char* name, int32_t row, int32_t col, float v) {
  FDEM_REQUIRE(m && name, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  int64_t lin;
  FDEM_TRY(cell_offset(m, row, col, &lin));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  FDEM_CUDA_TRY(cudaMemcpy(l->d + lin, &v, sizeof(float), cudaMemcpyHostToDevice));
  if (l->name == "obstacle") m->obstacle_full_clear = true;
  return FDEM_OK;
}

fdem_status fdem_map_clear(fdem_map* m, const char* name) {
  FDEM_REQUIRE(m && name, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  launch_fill(l->d, m->cells, kNaN, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  return FDEM_OK;
}

fdem_status fdem_map_clear_all(fdem_map* m) {
  FDEM_REQUIRE(m, "null map");
  DeviceGuard dg(m->device);
  for (auto& l : m->layers) launch_fill(l.d, m->cells, kNaN, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  return FDEM_OK;
}

fdem_status fdem_map_clear_at(fdem_map* m, int32_t row, int32_t col) {
  FDEM_REQUIRE(m, "null map");
  DeviceGuard dg(m->device);
  int64_t lin;
  FDEM_TRY(cell_offset(m, row, col, &lin));
  launch_clear_cell(layer_table(m), lin, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  return FDEM_OK;
}

fdem_status fdem_map_is_empty(fdem_map* m, int32_t* empty) {
  FDEM_REQUIRE(m && empty, "null argument");
  DeviceGuard dg(m->device);
  FDEM_CUDA_TRY(cudaMemsetAsync(m->d_flag + 1, 0, sizeof(uint32_t), m->stream));
  launch_any_not_nan(layer_ptr(m, "elevation"), m->cells, m->d_flag + 1, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  uint32_t flag = 0;
  FDEM_CUDA_TRY(cudaMemcpyAsync(&flag, m->d_flag + 1, sizeof(flag), cudaMemcpyDeviceToHost,
                                m->stream));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  *empty = flag ? 0 : 1;
  return FDEM_OK;
}

fdem_status fdem_map_sync(fdem_map* m) {
  FDEM_REQUIRE(m, "null map");
  DeviceGuard dg(m->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  return FDEM_OK;
}

void* fdem_map_stream(fdem_map* m) { return m ? m->stream : nullptr; }

// ───────────────────────────── mapper ────────────────────────────────────────

fdem_status fdem_mapper_create(fdem_map* map, const fdem_config* cfg, fdem_mapper** out) {
  FDEM_REQUIRE(map && out, "null argument");
  *out = nullptr;
  fdem_config c;
  if (cfg) c = *cfg; else fdem_config_default(&c);
  FDEM_TRY(validate_config(&c));
  DeviceGuard dg(map->device);
  fdem_mapper* mp = new (std::nothrow) fdem_mapper();
  if (!mp) return set_error(FDEM_ERR_OUT_OF_MEMORY, "host allocation failed");
  mp->map = map;
  mp->cfg = c;
  fdem_status st = ensure_estimator_layers(map, c);
  if (st != FDEM_OK) { delete mp; return st; }
  cudaError_t e = cudaMalloc(&mp->d_counters, CNT_COUNT * sizeof(uint32_t));
  if (e != cudaSuccess) { delete mp; return set_error(FDEM_ERR_CUDA, cudaGetErrorString(e)); }
  *out = mp;
  return FDEM_OK;
}

fdem_status fdem_mapper_destroy(fdem_mapper* mp) {
  if (!mp) return FDEM_OK;
  DeviceGuard dg(mp->map->device);
  cudaStreamSynchronize(mp->map->stream);
  free_scratch(mp);
  cudaFree(mp->d_counters);
  delete mp;
  return FDEM_OK;
}

fdem_status fdem_mapper_set_config(fdem_mapper* mp, const fdem_config* cfg) {
  FDEM_REQUIRE(mp && cfg, "null argument");
  FDEM_TRY(validate_config(cfg));
  DeviceGuard dg(mp->map->device);
  mp->cfg = *cfg;
  // setEstimatorType/setMappingMode re-create ElevationMapping (fastdem.cpp:28-38), whose
  // ctor adds whatever estimator layers are missing; existing data persists
  return ensure_estimator_layers(mp->map, mp->cfg);
}

fdem_status fdem_mapper_get_config(fdem_mapper* mp, fdem_config* out) {
  FDEM_REQUIRE(mp && out, "null argument");
  *out = mp->cfg;
  return FDEM_OK;
}

static fdem_status integrate_common(fdem_mapper* mp, const float* xyzw, const float* cov9,
                                    const float* intensity, const uint8_t* rgb, size_t n,
                                    const double* Tbs, const double* Twb) {
  FDEM_REQUIRE(mp, "null mapper");
  FDEM_REQUIRE(Tbs && Twb, "null transform");
  DeviceGuard dg(mp->map->device);
  if (n == 0) {
    // empty cloud: warn + false, nothing touched (fastdem.cpp:125-128)
    mp->last = fdem_scan_stats{};
    mp->last_had_work = false;
    return FDEM_OK;
  }
  FDEM_REQUIRE(xyzw, "xyzw is null");
  ScanInputs in{};
  in.xyzw = xyzw;
  in.intensity = intensity;
  in.rgb = rgb;
  in.cov9 = cov9;
  in.n = n;
  in.input_frame = INPUT_SENSOR_FRAME;
  in.Tbs = Tbs;
  in.Twb = Twb;
  in.robot_x = Twb[12];  // T_world_base.translation().head<2>() (fastdem.cpp:144)
  in.robot_y = Twb[13];
  return enqueue_scan(mp, in);
}

fdem_status fdem_mapper_integrate(fdem_mapper* mp, const float* xyzw, const float* intensity,
                                  const uint8_t* rgb, size_t n, const double* Tbs,
                                  const double* Twb, fdem_scan_stats* stats) {
  FDEM_TRY(integrate_common(mp, xyzw, nullptr, intensity, rgb, n, Tbs, Twb));
  DeviceGuard dg(mp->map->device);
  return finish_scan(mp, stats);
}

fdem_status fdem_mapper_integrate_with_cov(fdem_mapper* mp, const float* xyzw, const float* cov9,
                                           const float* intensity, const uint8_t* rgb, size_t n,
                                           const double* Tbs, const double* Twb,
                                           fdem_scan_stats* stats) {
  FDEM_REQUIRE(cov9 || n == 0, "cov9 is null");
  FDEM_TRY(integrate_common(mp, xyzw, cov9, intensity, rgb, n, Tbs, Twb));
  DeviceGuard dg(mp->map->device);
  return finish_scan(mp, stats);
}

fdem_status fdem_mapper_integrate_async(fdem_mapper* mp, const float* xyzw,
                                        const float* intensity, const uint8_t* rgb, size_t n,
                                        const double* Tbs, const double* Twb) {
  return integrate_common(mp, xyzw, nullptr, intensity, rgb, n, Tbs, Twb);
}

fdem_status fdem_mapper_wait(fdem_mapper* mp, fdem_scan_stats* stats) {
  FDEM_REQUIRE(mp, "null mapper");
  DeviceGuard dg(mp->map->device);
  return finish_scan(mp, stats);
}

fdem_status fdem_mapper_update(fdem_mapper* mp, const float* xyzw, const float* var_z,
                               const float* intensity, const uint8_t* rgb, size_t n,
                               double robot_x, double robot_y, fdem_scan_stats* stats) {
  FDEM_REQUIRE(mp, "null mapper");
  DeviceGuard dg(mp->map->device);
  if (n == 0) {
    // update() still moves a LOCAL map before rasterize() finds nothing
    // (elevation_mapping.cpp:111-117)
    mp->last = fdem_scan_stats{};
    mp->last_had_work = false;
    if (mp->cfg.mode == FDEM_MODE_LOCAL)
      FDEM_TRY(fdem_map_move(mp->map, robot_x, robot_y, mp->cfg.move_clear_policy, nullptr));
    if (stats) *stats = mp->last;
    return FDEM_OK;
  }
  FDEM_REQUIRE(xyzw, "xyzw is null");
  ScanInputs in{};
  in.xyzw = xyzw;
  in.intensity = intensity;
  in.rgb = rgb;
  in.var_z = var_z;
  in.n = n;
  in.input_frame = INPUT_MAP_FRAME;
  in.robot_x = robot_x;
  in.robot_y = robot_y;
  FDEM_TRY(enqueue_scan(mp, in));
  return finish_scan(mp, stats);
}

fdem_status fdem_mapper_last_preprocessed(fdem_mapper* mp, float* xyzw, float* cov9,
                                          int32_t* src_index, int64_t* n_kept) {
  FDEM_REQUIRE(mp && n_kept, "null argument");
  DeviceGuard dg(mp->map->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(mp->map->stream));
  *n_kept = 0;
  if (!mp->last_had_work || mp->last_n == 0) return FDEM_OK;
  if (cov9) return set_error(FDEM_ERR_UNSUPPORTED, "full covariances are not retained on device");
  // d_pm holds the map-frame points in input order, dropped points marked NaN; compact on
  // the host (observation hook, not the hot path)
  std::vector<float> pm(static_cast<size_t>(mp->last_n) * 4);
  FDEM_CUDA_TRY(cudaMemcpy(pm.data(), mp->d_pm, pm.size() * sizeof(float), cudaMemcpyDeviceToHost));
  int64_t k = 0;
  for (uint32_t i = 0; i < mp->last_n; ++i) {
    if (std::isnan(pm[4 * i])) continue;
    if (xyzw) {
      xyzw[4 * k + 0] = pm[4 * i + 0];
      xyzw[4 * k + 1] = pm[4 * i + 1];
      xyzw[4 * k + 2] = pm[4 * i + 2];
      xyzw[4 * k + 3] = pm[4 * i + 3];  // NOTE: w carries sigma_z^2 = cov(2,2), not 1
    }
    if (src_index) src_index[k] = static_cast<int32_t>(i);
    ++k;
  }
  *n_kept = k;
  return FDEM_OK;
}

fdem_status fdem_mapper_last_rasterized(fdem_mapper* mp, float* xyz, int64_t* n_cells) {
  FDEM_REQUIRE(mp && n_cells, "null argument");
  fdem_map* m = mp->map;
  DeviceGuard dg(m->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  *n_cells = 0;
  if (!mp->last_had_work || mp->last.n_cells == 0) return FDEM_OK;
  FDEM_TRY(refresh_geometry(m));
  uint32_t tc = 0;
  FDEM_TRY(current_touched_count(m, &tc));
  std::vector<uint32_t> keys(tc);
  std::vector<float> minz(tc);
  FDEM_CUDA_TRY(cudaMemcpy(keys.data(), m->d_touched_keys, tc * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  FDEM_CUDA_TRY(cudaMemcpy(minz.data(), m->d_touched_minz, tc * sizeof(float), cudaMemcpyDeviceToHost));
  const int rows_local = m->geom.row_end - m->geom.row_begin;
  int64_t k = 0;
  for (uint32_t i = 0; i < tc; ++i) {
    if (keys[i] == static_cast<uint32_t>(m->cells)) continue;
    if (xyz) {
      const int32_t col = static_cast<int32_t>(keys[i] / rows_local);
      const int32_t row = static_cast<int32_t>(keys[i] % rows_local) + m->geom.row_begin;
      double x, y;
      geom_cell_position(m->geom, row, col, x, y);
      xyz[3 * k + 0] = static_cast<float>(x);  // Vector3f(pos.x(), pos.y(), min_z), fastdem.cpp:210
      xyz[3 * k + 1] = static_cast<float>(y);
      xyz[3 * k + 2] = minz[i];
    }
    ++k;
  }
  *n_cells = k;
  return FDEM_OK;
}

fdem_status fdem_mapper_launch_count(fdem_mapper* mp, int64_t* launches) {
  FDEM_REQUIRE(mp && launches, "null argument");
  *launches = mp->map->lc.mine;
  return FDEM_OK;
}

}  // extern "C"

// raycasting / voxel / inpainting entry points live in capi_post.cu
#include "capi_post.inc"
