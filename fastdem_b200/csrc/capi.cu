// capi.cu — implementation of include/fastdem_b200.h: device-resident ElevationMap,
// FastDEM mapper, and the per-scan kernel pipeline.  Host logic only; all arithmetic on
// map state happens in kernels.cu / kernels_raycast.cu.  There is no CPU fallback: every
// entry point needs a working CUDA device.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

// a FastDEM whose map was destroyed under it (fdem_map_destroy detaches it)
#define FDEM_MAPPER_ALIVE(mp)                                                               \
  do {                                                                                      \
    if ((mp) && !(mp)->map)                                                                 \
      return set_error(FDEM_ERR_INVALID_ARGUMENT, "the mapper's map has been destroyed");   \
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
  // Layers the reference creates lazily on a condition only the device learns (first scan WITH
  // observations that carries the channel; first applyRaycasting that passes its guards) are
  // allocated when first needed but stay invisible to the map API until the device has reported
  // the StateFlag that reveals them — exists() / getLayers() then match the reference's.
  uint32_t hidden_until = 0;  // 0 = visible
};

// what the last kernel of a scan leaves for the host (copied D2H into pinned memory)
struct ScanResult {
  uint32_t counters[CNT_COUNT];
  DeviceState state;
};

}  // namespace

constexpr int kResultRing = 8;   // scans that may be in flight before a result slot is reused
constexpr int kStageRing = 3;    // host-input staging buffers: the copies of scans k+1 and k+2 can be queued
                                 // while scan k runs, so the copy engine never waits for the host's next submit

struct fdem_map {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  GridGeom geom{};            // host mirror; valid when !geom_stale
  bool geom_stale = false;    // async scans in flight may have moved the window
  DeviceState* d_state = nullptr;  // [2]
  size_t cells = 0;           // rows_local * cols
  std::vector<Layer> layers;
  // touched-cell list of the last observing scan (obstacle reset)
  uint32_t* d_touched_keys = nullptr;
  float* d_touched_minz = nullptr;
  size_t touched_cap = 0;
  bool obstacle_full_clear = false;  // obstacle was written behind the mapper's back: the next
                                     // scan sets SF_OBSTACLE_DIRTY on the device first
  uint32_t state_flags = 0;          // host mirror of DeviceState::flags (valid when !geom_stale)
  std::vector<fdem_mapper*> mappers; // FastDEM objects bound to this map (they hold map-sized scratch)
  // raycasting scratch (allocated on first use)
  uint32_t* d_ray_min_enc = nullptr;
  uint32_t* d_hits = nullptr;
  // small scratch
  uint32_t* d_flag = nullptr;
  // ring of result slots in pinned + mapped host memory: the kernel that ends scan q writes
  // slot q % kResultRing directly; ev_scan[q % kResultRing] marks that scan's completion
  ScanResult* h_result = nullptr;
  uint32_t* d_result_host = nullptr;  // device-side alias of h_result[0]
  cudaEvent_t ev_scan[kResultRing] = {};
  uint64_t seq = 0;                   // scans enqueued so far (ticket of the next scan)
  uint64_t layer_epoch = 0;           // bumped whenever a layer is added (lookup caches)
  LaunchCounter lc;
  // cell-sized scratch layers for the post-process stencils (snapshots / double buffers),
  // allocated on first use and kept
  float* d_tmp[2] = {nullptr, nullptr};
  // last fdem_map_pack_pointcloud2() result (device resident)
  float* d_pack = nullptr;
  size_t pack_cap_bytes = 0;
  uint32_t* d_pack_cols = nullptr;    // [2 * cols + 1] per-column counts, offsets, total
  std::vector<std::string> pack_fields;
  uint32_t pack_width = 0;
};

namespace {
// one scan's kernel arguments, kept alive in the mapper so graph nodes can be patched in place
struct ScanLaunch {
  PreprocessParams pp;
  CommitParams cp;
  ScatterParams sp;
  EstimateParams ep;
  LayerTable lt;
  TileBuffers tb;
  const DeviceState* st_cur;
  DeviceState* st_next;
  uint32_t* counters;
  float4* pm;
  uint32_t* keys;
  uint32_t* vals;
  uint32_t* host_out;
  PublishArgs pub;
};
// GN_K3L (the warp-per-bucket pass of K3t) is part of the chain only when the scan shape is
// "light" (TileBuffers::job_counter == CNT_HEAVY)
enum { GN_K1 = 0, GN_K2, GN_SCATTER, GN_K3L, GN_K3T, GN_COUNT };
struct ScanGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t node[GN_COUNT] = {};
  ScanLaunch cached{};
  bool valid = false;
};
}  // namespace

struct fdem_mapper {
  fdem_map* map = nullptr;   // null once the map has been destroyed under the mapper
  int device = 0;
  fdem_config cfg{};
  size_t cap = 0;  // scratch capacity in points
  // host-input staging, double buffered: slot q % kStageRing of scan q
  float4* d_in_xyzw[kStageRing] = {};
  float* d_in_intensity[kStageRing] = {};
  uint8_t* d_in_rgb[kStageRing] = {};
  float* d_in_aux[kStageRing] = {};  // cov9 (N x 9) or var_z (N)
  uint8_t* d_in_raw[kStageRing] = {};  // PointCloud2 message bodies
  size_t raw_cap = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[kStageRing] = {};
  uint64_t last_ticket = 0;
  float4* d_pm = nullptr;
  uint32_t *d_keys = nullptr, *d_vals = nullptr, *d_skeys = nullptr, *d_svals = nullptr;
  uint64_t *d_vkeys = nullptr, *d_svkeys = nullptr;  // 63-bit voxel keys (raycasting, unbounded range)
  // raycasting branch (runs beside scatter + K3t on aux_stream): its own key / value buffers
  uint32_t *d_vk32 = nullptr, *d_svk32 = nullptr, *d_vv = nullptr, *d_svv = nullptr;
  float4* d_rays = nullptr;                          // rays to trace (end points), (length, azimuth) bundles
  float4* d_rays_tmp = nullptr;                      // the same in discovery order
  uint32_t* d_ray_hist = nullptr;                    // counting-sort histogram + cursors
  uint32_t* d_vox_rows = nullptr;                    // MSD voxel sort: row tables (voxel_rows_scratch_words)
  uint32_t vox_rows_cap = 0;
  bool voxel_sort_library = true;                    // false: the 2-level MSD sort of kernels_raycast.cu
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_geom = nullptr, ev_rays = nullptr;
  void* d_sort_temp = nullptr;
  size_t sort_temp_bytes = 0;
  uint32_t* d_counters = nullptr;
  fdem_scan_stats last{};
  uint32_t last_n = 0;
  bool last_raw = false;
  bool last_had_work = false;
  bool pending = false;  // async scans queued since the last wait
  // tile path (2-level sort-by-cell); global CUB sort kept as the alternative path
  bool use_tile = true;
  bool use_graph = true;   // launch the tile pipeline as one CUDA graph (FDEM_GRAPH=0 disables)
  bool use_pdl = false;    // programmatic edges between the graph's kernels (FDEM_PDL=1 enables)
  bool counters_dirty = true;
  ScanLaunch launch{};
  ScanGraph sg;
  bool tile_dirty = true;  // L1 scratch must be zeroed before the next scan
  TileBuffers tb{};
  uint32_t max_buckets = 0;   // bucket arrays are sized for the smallest bucket shape
  bool bucket_bits_auto = true;  // pick the K3t bucket shape from the last scan's density
  bool tile_light_auto = true;   // ... and whether the warp-per-bucket pass runs first
  bool tile_light_pin = false;   // FDEM_TILE_LIGHT=1: always, where the bucket shape allows it
  // batched integration (fdem_mapper_integrate_batch): a second scratch set so scan k+1's
  // front half can run beside scan k's estimator, a ring of state slots, the batch graph
  float4* d_pm2 = nullptr;
  uint32_t* d_keys2 = nullptr;
  TileBuffers tb2{};
  size_t cap2 = 0;
  uint32_t* d_batch_counters = nullptr;  // [16][CNT_COUNT]: one counter block per scan of a batch
  uint32_t* h_batch_counters = nullptr;  // pinned copy, written by a memcpy node at the end of every batch
  int last_batch_n = 0;                  // scans in the most recent batch (fdem_mapper_last_batch_stats)
  uint64_t last_batch_ticket0 = 0;
  size_t last_batch_points[16] = {};
  DeviceState* d_ring = nullptr;   // [kMaxBatch + 1]
  MoveRecord* d_move = nullptr;    // [2]
  struct BatchGraphState* bg = nullptr;   // [2]: executable graphs used alternately, so one can be
                                          // re-parameterised while the other is still executing
  int bg_flip = 0;
  // by-name layer lookups of one scan (~25 string searches), cached until a layer is added
  uint64_t cached_epoch = ~0ull;
  EstLayers cached_L{};
  LayerTable cached_lt{};
  // optional per-stage device timing (bench / profiling)
  bool stage_timing = false;
  std::vector<cudaEvent_t> ev_pool;               // free events
  std::vector<std::vector<cudaEvent_t>> ev_scans;  // FDEM_STAGE_COUNT+1 events per in-flight scan
  double stage_ms[FDEM_STAGE_COUNT] = {0, 0, 0, 0, 0, 0};
  int64_t stage_scans = 0;
};

namespace {

struct DeviceGuard {
  int prev = 0;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    else (void)cudaGetLastError();  // a bad ordinal must not linger as the next call's "last error"
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

void apply_state_flags(fdem_map* m, uint32_t flags) {
  m->state_flags = flags;
  for (auto& l : m->layers)
    if (l.hidden_until & flags) l.hidden_until = 0;
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
    ++m->layer_epoch;
    l = &m->layers.back();
  }
  launch_fill(l->d, m->cells, fill, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  if (l->name == "obstacle") m->obstacle_full_clear = true;
  return FDEM_OK;
}

fdem_status ensure_layer(fdem_map* m, const char* name, float fill) {
  if (Layer* l = find_layer(m, name)) {
    l->hidden_until = 0;
    return FDEM_OK;
  }
  return add_layer(m, name, fill);
}

// storage for a lazily created layer: invisible until the device reports `flag`
fdem_status ensure_hidden_layer(fdem_map* m, const char* name, float fill, uint32_t flag) {
  if (find_layer(m, name)) return FDEM_OK;
  FDEM_TRY(add_layer(m, name, fill));
  if (!(m->state_flags & flag)) m->layers.back().hidden_until = flag;
  return FDEM_OK;
}

// bring the host mirror of the geometry up to date with the device
fdem_status refresh_geometry(fdem_map* m) {
  if (!m->geom_stale) return FDEM_OK;
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  DeviceState st;
  FDEM_CUDA_TRY(cudaMemcpy(&st, m->d_state, sizeof(st), cudaMemcpyDeviceToHost));
  m->geom = st.geom;
  m->geom_stale = false;
  apply_state_flags(m, st.flags);
  return FDEM_OK;
}

// what the map API sees: hidden layers do not exist (yet).  Scans still in flight may reveal
// one, so their results are awaited first when it matters.
Layer* find_visible_layer(fdem_map* m, const char* name) {
  Layer* l = find_layer(m, name);
  if (l && l->hidden_until && m->geom_stale) (void)refresh_geometry(m);
  return (l && !l->hidden_until) ? l : nullptr;
}
void settle_hidden_layers(fdem_map* m) {
  if (!m->geom_stale) return;
  for (auto& l : m->layers)
    if (l.hidden_until) { (void)refresh_geometry(m); return; }
}

fdem_status push_state(fdem_map* m, uint32_t touched_count) {
  DeviceState st{};
  st.geom = m->geom;
  st.touched_count = touched_count;
  st.flags = m->state_flags;
  FDEM_CUDA_TRY(cudaMemcpyAsync(m->d_state, &st, sizeof(st), cudaMemcpyHostToDevice, m->stream));
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
  for (int i = 0; i < kStageRing; ++i) {
    cudaFree(mp->d_in_xyzw[i]);
    cudaFree(mp->d_in_intensity[i]);
    cudaFree(mp->d_in_rgb[i]);
    cudaFree(mp->d_in_aux[i]);
    mp->d_in_xyzw[i] = nullptr; mp->d_in_intensity[i] = nullptr;
    mp->d_in_rgb[i] = nullptr; mp->d_in_aux[i] = nullptr;
  }
  cudaFree(mp->d_pm);
  cudaFree(mp->d_keys);
  cudaFree(mp->d_vals);
  cudaFree(mp->d_skeys);
  cudaFree(mp->d_svals);
  cudaFree(mp->d_vkeys);
  cudaFree(mp->d_svkeys);
  cudaFree(mp->d_vk32);
  cudaFree(mp->d_svk32);
  cudaFree(mp->d_vv);
  cudaFree(mp->d_svv);
  cudaFree(mp->d_rays);
  cudaFree(mp->d_rays_tmp);
  cudaFree(mp->d_ray_hist);
  cudaFree(mp->d_vox_rows);
  mp->d_vox_rows = nullptr;
  mp->vox_rows_cap = 0;
  mp->d_vk32 = mp->d_svk32 = mp->d_vv = mp->d_svv = nullptr;
  mp->d_rays = mp->d_rays_tmp = nullptr;
  mp->d_ray_hist = nullptr;
  cudaFree(mp->d_sort_temp);
  for (int i = 0; i < kStageRing; ++i) { cudaFree(mp->d_in_raw[i]); mp->d_in_raw[i] = nullptr; }
  mp->raw_cap = 0;
  cudaFree(mp->tb.records);
  mp->tb.records = nullptr;
  mp->d_pm = nullptr; mp->d_keys = mp->d_vals = nullptr;
  mp->d_skeys = mp->d_svals = nullptr; mp->d_vkeys = mp->d_svkeys = nullptr;
  mp->d_sort_temp = nullptr;
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
  if (mp->copy_stream) FDEM_CUDA_TRY(cudaStreamSynchronize(mp->copy_stream));
  if (mp->aux_stream) FDEM_CUDA_TRY(cudaStreamSynchronize(mp->aux_stream));
  const size_t cap = std::max<size_t>(std::max(n, mp->cap), 1024);
  free_scratch(mp);
  for (int i = 0; i < kStageRing; ++i) {
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_xyzw[i], cap * sizeof(float4)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_intensity[i], cap * sizeof(float)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_rgb[i], cap * 3 + 16));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_aux[i], cap * 9 * sizeof(float)));
  }
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_pm, cap * sizeof(float4)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_keys, cap * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_vals, cap * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_skeys, cap * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_svals, cap * sizeof(uint32_t)));
  size_t temp = sort_pairs_u32_temp_bytes(static_cast<uint32_t>(cap), 32);
  if (need_vox) {
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_vkeys, cap * sizeof(uint64_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_svkeys, cap * sizeof(uint64_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_vk32, cap * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_svk32, cap * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_vv, cap * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_svv, cap * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_rays, cap * sizeof(float4)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_rays_tmp, cap * sizeof(float4)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_ray_hist, ray_sort_scratch_words() * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMemset(mp->d_ray_hist, 0, ray_sort_scratch_words() * sizeof(uint32_t)));
    temp = std::max(temp, sort_pairs_u64_temp_bytes(static_cast<uint32_t>(cap), 63));
  }
  FDEM_CUDA_TRY(cudaMalloc(&mp->d_sort_temp, temp));
  mp->sort_temp_bytes = temp;
  FDEM_CUDA_TRY(cudaMalloc(&mp->tb.records, cap * sizeof(CellRecord)));
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

fdem_status ensure_raycast_storage(fdem_map* m, bool hidden);
__global__ void or_state_flags_kernel(DeviceState* st, uint32_t bits) { st->flags |= bits; }
RaycastParams raycast_params(fdem_map* m, const fdem_config& cfg, const float origin[3]);
bool voxel_box(const fdem_config& cfg, const float* T2, float voxel, VoxelBox* box);
fdem_status launch_scan_graph(fdem_mapper* mp, cudaStream_t s);

cudaEvent_t take_event(fdem_mapper* mp) {
  if (!mp->ev_pool.empty()) {
    cudaEvent_t e = mp->ev_pool.back();
    mp->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// fold finished scans' stage events into the accumulators (stream must be idle)
void drain_stage_events(fdem_mapper* mp) {
  for (auto& evs : mp->ev_scans) {
    for (int st = 0; st < FDEM_STAGE_COUNT; ++st) {
      float ms = 0.0f;
      if (cudaEventElapsedTime(&ms, evs[st], evs[st + 1]) == cudaSuccess) mp->stage_ms[st] += ms;
    }
    ++mp->stage_scans;
    for (cudaEvent_t e : evs) mp->ev_pool.push_back(e);
  }
  mp->ev_scans.clear();
  (void)cudaGetLastError();
}

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
  const uint8_t* raw = nullptr;                    // PointCloud2 body (xyzw etc. null then)
  const fdem_pointcloud2_layout* layout = nullptr;
  bool known_device = false;                       // the caller has already checked: device pointers
};

// build_only: fill *build_only with the scan's kernel arguments (against the primary scratch
// set, for result slot `ticket_override`) and return without launching anything
fdem_status enqueue_scan(fdem_mapper* mp, const ScanInputs& in, ScanLaunch* build_only = nullptr,
                         uint64_t ticket_override = 0) {
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
  std::vector<cudaEvent_t>* evs = nullptr;
  if (mp->stage_timing) {
    mp->ev_scans.emplace_back();
    evs = &mp->ev_scans.back();
    for (int i = 0; i <= FDEM_STAGE_COUNT; ++i) evs->push_back(take_event(mp));
  }
  auto mark = [&](int boundary) {
    if (evs) cudaEventRecord((*evs)[boundary], s);
  };
  mark(FDEM_STAGE_H2D);
  // layers the cloud's channels need (updateIntensity / updateColor add them lazily,
  // elevation_mapping.cpp:155,169)
  const bool has_intensity = in.intensity || (in.layout && in.layout->off_intensity >= 0);
  const bool has_color = in.rgb || (in.layout && in.layout->off_rgb >= 0);
  if (has_intensity) FDEM_TRY(ensure_hidden_layer(m, "intensity", kNaN, SF_INTENSITY));
  if (has_color) FDEM_TRY(ensure_hidden_layer(m, "color", kNaN, SF_COLOR));
  if (cfg.raycasting_enabled && in.input_frame == INPUT_SENSOR_FRAME) {
    FDEM_REQUIRE(m->geom.row_begin == 0 && m->geom.row_end == m->geom.rows,
                 "raycasting is not supported on a row stripe");
    FDEM_TRY(ensure_raycast_storage(m, /*hidden=*/true));
  }

  PreprocessParams pp{};
  // Inputs already on the device are used in place.  Host inputs are copied into staging
  // slot (ticket % kStageRing) on a separate copy stream, so the copy of scan k+1 overlaps
  // the kernels of scan k; the compute stream waits for the copy through an event.
  FDEM_REQUIRE(!(in.cov9 && in.var_z), "cov9 and var_z are exclusive");
  const uint64_t ticket = build_only ? ticket_override : m->seq;
  const int slot = static_cast<int>(ticket % kStageRing);
  const bool overlap = !mp->stage_timing;  // stage timing wants the copy on the timed stream
  cudaStream_t cs = overlap ? mp->copy_stream : s;
  bool staged = false;
  auto stage_in = [&](const void* src, void* scratch, size_t bytes, const void** out) -> fdem_status {
    if (!src) { *out = nullptr; return FDEM_OK; }
    if (in.known_device || is_device_pointer(src)) { *out = src; return FDEM_OK; }
    if (!staged && overlap && ticket >= kStageRing) {
      // the slot's previous user (scan ticket - kStageRing) must be done with the buffers
      FDEM_CUDA_TRY(cudaStreamWaitEvent(cs, m->ev_scan[(ticket - kStageRing) % kResultRing], 0));
    }
    staged = true;
    FDEM_CUDA_TRY(cudaMemcpyAsync(scratch, src, bytes, cudaMemcpyHostToDevice, cs));
    *out = scratch;
    return FDEM_OK;
  };
  const void *xyzw_v = nullptr, *inten_v = nullptr, *rgb_v = nullptr, *aux_v = nullptr;
  const void* raw_v = nullptr;
  if (in.raw) {
    const fdem_pointcloud2_layout& lo = *in.layout;
    const size_t bytes = static_cast<size_t>(n) * lo.point_step;
    if (!is_device_pointer(in.raw) && mp->raw_cap < bytes) {
      FDEM_CUDA_TRY(cudaStreamSynchronize(s));
      FDEM_CUDA_TRY(cudaStreamSynchronize(mp->copy_stream));
      for (int i = 0; i < kStageRing; ++i) {
        cudaFree(mp->d_in_raw[i]);
        mp->d_in_raw[i] = nullptr;
        FDEM_CUDA_TRY(cudaMalloc(&mp->d_in_raw[i], bytes + 16));
      }
      mp->raw_cap = bytes;
    }
    FDEM_TRY(stage_in(in.raw, mp->d_in_raw[slot], bytes, &raw_v));
    FDEM_REQUIRE((reinterpret_cast<uintptr_t>(raw_v) & 3) == 0, "PointCloud2 data must be 4-byte aligned");
    pp.raw = static_cast<const uint8_t*>(raw_v);
    pp.point_step = lo.point_step;
    pp.off_x = lo.off_x; pp.off_y = lo.off_y; pp.off_z = lo.off_z;
    pp.off_intensity = lo.off_intensity; pp.intensity_type = lo.intensity_type; pp.off_rgb = lo.off_rgb;
    // K1 unpacks the channels into the staging slot's SoA buffers for the later kernels
    if (lo.off_intensity >= 0) { pp.out_intensity = mp->d_in_intensity[slot]; inten_v = pp.out_intensity; }
    if (lo.off_rgb >= 0) { pp.out_rgb = mp->d_in_rgb[slot]; rgb_v = pp.out_rgb; }
    // exactly one 4-byte channel in the fourth word (float32 intensity or packed rgb)
    const bool one_channel = (lo.off_intensity == 12 && lo.intensity_type == 7 && lo.off_rgb < 0) ||
                             (lo.off_rgb == 12 && lo.off_intensity < 0) ||
                             (lo.off_intensity < 0 && lo.off_rgb < 0);
    pp.raw_vec16 = (lo.point_step == 16 && lo.off_x == 0 && lo.off_y == 4 && lo.off_z == 8 && one_channel &&
                    (reinterpret_cast<uintptr_t>(raw_v) & 15) == 0) ? 1 : 0;
  }
  FDEM_TRY(stage_in(in.xyzw, mp->d_in_xyzw[slot], static_cast<size_t>(n) * 16, &xyzw_v));
  if (!in.raw) {
    FDEM_TRY(stage_in(in.intensity, mp->d_in_intensity[slot], static_cast<size_t>(n) * 4, &inten_v));
    FDEM_TRY(stage_in(in.rgb, mp->d_in_rgb[slot], static_cast<size_t>(n) * 3, &rgb_v));
  }
  if (in.cov9) FDEM_TRY(stage_in(in.cov9, mp->d_in_aux[slot], static_cast<size_t>(n) * 36, &aux_v));
  if (in.var_z) FDEM_TRY(stage_in(in.var_z, mp->d_in_aux[slot], static_cast<size_t>(n) * 4, &aux_v));
  if (staged && overlap) {
    FDEM_CUDA_TRY(cudaEventRecord(mp->ev_copied[slot], cs));
    FDEM_CUDA_TRY(cudaStreamWaitEvent(s, mp->ev_copied[slot], 0));
  }
  const float* xyzw_d = static_cast<const float*>(xyzw_v);
  const float* inten_d = static_cast<const float*>(inten_v);
  const uint8_t* rgb_d = static_cast<const uint8_t*>(rgb_v);
  const float* aux_d = static_cast<const float*>(aux_v);
  FDEM_REQUIRE((reinterpret_cast<uintptr_t>(xyzw_d) & 15) == 0, "xyzw must be 16-byte aligned");
  pp.xyzw = reinterpret_cast<const float4*>(xyzw_d);
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
  const bool tile = mp->use_tile;
  pp.bucket_count = tile ? mp->tb.bucket_count : nullptr;
  pp.bucket_bits = mp->tb.bucket_bits;
  pp.write_vals = tile ? 0 : 1;
  if (tile && mp->tile_dirty) {
    // first scan / after a failed scan: the L1 scratch K3t normally re-arms must be zero
    FDEM_CUDA_TRY(cudaMemsetAsync(mp->tb.bucket_count, 0, mp->max_buckets * sizeof(uint32_t), s));
    FDEM_CUDA_TRY(cudaMemsetAsync(mp->tb.bucket_cursor, 0, mp->max_buckets * sizeof(uint32_t), s));
    mp->tile_dirty = false;
  }

  const DeviceState* st_in = m->d_state;     // current (committed) state
  DeviceState* st_out = m->d_state + 1;      // this scan's state; publish makes it current

  if (mp->counters_dirty) {
    // first scan / after a failed scan; afterwards the publish kernel re-arms the counters
    FDEM_CUDA_TRY(cudaMemsetAsync(mp->d_counters, 0, CNT_COUNT * sizeof(uint32_t), s));
    mp->counters_dirty = false;
  }
  if (m->obstacle_full_clear && !build_only) {
    // obstacle was uploaded / edited by the caller: the next scan WITH observations clears the
    // whole layer like the reference (elevation_mapping.cpp:146) instead of last scan's cells.
    // Decided on the device (a scan without observations must leave the layer alone, :116-117).
    or_state_flags_kernel<<<1, 1, 0, s>>>(m->d_state, SF_OBSTACLE_DIRTY);
    ++m->lc.mine;
    m->obstacle_full_clear = false;
  }

  ScanLaunch& L = mp->launch;
  L.pp = pp;
  L.cp = CommitParams{};
  L.cp.robot_x = in.robot_x;
  L.cp.robot_y = in.robot_y;
  L.cp.local_mode = pp.local_mode;
  L.cp.clear_policy = cfg.move_clear_policy;
  L.cp.invalid_key = pp.invalid_key;
  L.cp.obstacle = nullptr;  // set below from the cached layer lookup
  L.cp.touched_keys = m->d_touched_keys;
  L.cp.tile_path = tile ? 1 : 0;
  L.cp.tb = mp->tb;
  L.cp.flags_if_cells = (has_intensity ? SF_INTENSITY : 0u) | (has_color ? SF_COLOR : 0u);
  L.cp.raycast = (cfg.raycasting_enabled && in.input_frame == INPUT_SENSOR_FRAME) ? 1 : 0;
  if (L.cp.raycast) {
    // applyRaycasting's isInside guard tests the float sensor origin widened to double (raycasting.cpp:230)
    L.cp.rc_origin_x = static_cast<double>(static_cast<float>(T[12]));
    L.cp.rc_origin_y = static_cast<double>(static_cast<float>(T[13]));
  }
  L.cp.obstacle_cells = m->cells;
  L.sp = ScatterParams{};
  L.sp.keys = mp->d_keys;
  L.sp.pm = mp->d_pm;
  L.sp.intensity = inten_d;
  L.sp.n = n;
  L.sp.invalid_key = pp.invalid_key;
  L.sp.tb = mp->tb;
  L.sp.counters = mp->d_counters;
  L.sp.st_cur = m->d_state;
  L.sp.obstacle = L.cp.obstacle;
  L.sp.touched_keys = m->d_touched_keys;
  L.sp.obstacle_cells = m->cells;
  EstimateParams& ep = L.ep;
  ep = EstimateParams{};
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
  if (mp->cached_epoch != m->layer_epoch) {
    mp->cached_L = est_layers(m);
    mp->cached_lt = layer_table(m);
    mp->cached_epoch = m->layer_epoch;
  }
  ep.L = mp->cached_L;
  L.lt = mp->cached_lt;
  L.cp.obstacle = ep.L.obstacle;
  L.sp.obstacle = ep.L.obstacle;
  L.tb = mp->tb;
  L.st_cur = m->d_state;
  L.st_next = m->d_state + 1;
  L.counters = mp->d_counters;
  L.pm = mp->d_pm;
  L.keys = mp->d_keys;
  L.vals = mp->d_vals;
  uint32_t* host_slot = m->d_result_host + (ticket % kResultRing) * (sizeof(ScanResult) / 4);
  L.host_out = host_slot;

  const bool raycast = cfg.raycasting_enabled && in.input_frame == INPUT_SENSOR_FRAME;
  // K3t's last CTA ends the scan itself unless more kernels follow (raycasting)
  L.pub = PublishArgs{};
  L.pub.enabled = (tile && !raycast) ? 1 : 0;
  L.pub.st_cur = m->d_state;
  L.pub.host_out = host_slot;
  if (build_only) {
    *build_only = L;
    return FDEM_OK;
  }
  const bool graph = tile && mp->use_graph && !mp->stage_timing && !raycast;
  fdem_status launch_status = FDEM_OK;
  if (graph) {
    // K1 -> K2 -> scatter -> K3t -> publish as one cudaGraphLaunch; only the nodes whose
    // arguments changed since the last scan are patched
    launch_status = launch_scan_graph(mp, s);
    m->lc.mine += mp->tb.job_counter == CNT_HEAVY ? GN_COUNT : GN_COUNT - 1;
  } else {
    mark(FDEM_STAGE_PREPROCESS);
    launch_preprocess_bin(pp, st_in, mp->d_counters, mp->d_pm, mp->d_keys, mp->d_vals, s, m->lc);
    mark(FDEM_STAGE_COMMIT);
    launch_commit(L.cp, st_in, st_out, mp->d_counters, L.lt, s, m->lc);
    // ── raycasting branch (fastdem.cpp:153-159), queued BESIDE the cell reduction: voxelGrid(ANY)
    // and the DDA read only K1's map-frame points and the committed geometry and write only
    // scan scratch, so they run on the aux stream while scatter + K3t update the map; the
    // branches join at the resolve kernel, the first raycasting step that touches a layer.
    // (With stage timing on — or on the global-sort path, which shares the CUB scratch — the
    // branch is issued on the main stream in its own stage instead.) ──
    const bool rc_beside = raycast && tile && !mp->stage_timing;
    cudaStream_t aux = rc_beside ? mp->aux_stream : s;
    RaycastParams rcp{};
    auto raycast_branch = [&]() -> fdem_status {
      // sensor origin = (T_world_base*T_base_sensor).translation(), ray_scan = voxelGrid(points, resolution, ANY)
      const float origin[3] = {static_cast<float>(T[12]), static_cast<float>(T[13]),
                               static_cast<float>(T[14])};
      const float voxel = static_cast<float>(m->geom.res);
      if (voxel < 0.001f || voxel > 100.0f)
        return set_error(FDEM_ERR_INVALID_ARGUMENT, "voxel_size must be in [0.001, 100]");
      rcp = raycast_params(m, cfg, origin);
      if (rc_beside) {
        FDEM_CUDA_TRY(cudaEventRecord(mp->ev_geom, s));
        FDEM_CUDA_TRY(cudaStreamWaitEvent(aux, mp->ev_geom, 0));
      }
      // voxelGrid(ANY): sort by voxel key.  The crop filters usually bound the kept points
      // to a box small enough for 32-bit keys (4 radix passes instead of 8).
      VoxelBox box{};
      cudaError_t se = cudaSuccess;
      const RaySortScratch rss{mp->d_ray_hist, mp->d_rays_tmp, mp->d_rays};
      const bool have_box = voxel_box(cfg, pp.T2, voxel, &box);
      if (have_box && !mp->voxel_sort_library && box.by + box.bz <= 22) {
        // our own 2-level MSD sort (rows = (z, y), then x on chip): no library launches
        const uint32_t n_rows = 1u << (box.by + box.bz);
        if (mp->vox_rows_cap < n_rows) {
          FDEM_CUDA_TRY(cudaStreamSynchronize(aux));
          cudaFree(mp->d_vox_rows);
          mp->d_vox_rows = nullptr;
          mp->vox_rows_cap = 0;
          const size_t words = voxel_rows_scratch_words(n_rows);
          FDEM_CUDA_TRY(cudaMalloc(&mp->d_vox_rows, words * sizeof(uint32_t)));
          FDEM_CUDA_TRY(cudaMemset(mp->d_vox_rows, 0, words * sizeof(uint32_t)));
          mp->vox_rows_cap = n_rows;
        }
        const VoxelRowsScratch vs{mp->d_vox_rows, mp->vox_rows_cap, mp->d_vk32,
                                  reinterpret_cast<unsigned long long*>(mp->d_vkeys)};
        launch_voxel_select_rays_msd(mp->d_pm, n, 1.0f / voxel, box, vs, rcp, st_out, mp->d_counters, rss, aux, m->lc);
      } else if (have_box) {
        const int bits = box.bx + box.by + box.bz + 1;
        launch_voxel_keys32(mp->d_pm, n, 1.0f / voxel, box, mp->d_vk32, mp->d_vv, mp->d_counters, aux, m->lc);
        se = sort_pairs_u32(mp->d_sort_temp, mp->sort_temp_bytes, mp->d_vk32, mp->d_svk32,
                            mp->d_vv, mp->d_svv, n, bits, aux, m->lc);
        launch_voxel_select_rays32(mp->d_svk32, mp->d_svv, n, box.invalid_key, rcp, st_out, mp->d_pm,
                                   mp->d_counters, rss, aux, m->lc);
      } else {
        launch_voxel_keys(mp->d_pm, n, 1.0f / voxel, mp->d_vkeys, mp->d_vv, aux, m->lc);
        se = sort_pairs_u64(mp->d_sort_temp, mp->sort_temp_bytes, mp->d_vkeys, mp->d_svkeys,
                            mp->d_vv, mp->d_svv, n, 64, aux, m->lc);
        launch_voxel_select_rays64(mp->d_svkeys, mp->d_svv, n, rcp, st_out, mp->d_pm, mp->d_counters,
                                   rss, aux, m->lc);
      }
      if (se != cudaSuccess) return set_error(FDEM_ERR_CUDA, cudaGetErrorString(se));
      launch_raycast_dda(rcp, st_out, mp->d_rays, n, mp->d_counters, aux, m->lc);
      if (rc_beside) FDEM_CUDA_TRY(cudaEventRecord(mp->ev_rays, aux));
      return FDEM_OK;
    };
    if (rc_beside) launch_status = raycast_branch();
    mark(FDEM_STAGE_SORT);
    if (tile) {
      // L1 of the 2-level sort: one scatter pass into the bucket segments K2 allocated
      launch_scatter_records(L.sp, s, m->lc);
    } else {
      const int bits = key_bits(m->cells);  // keys are in [0, cells]; `cells` = dropped point
      FDEM_CUDA_TRY(sort_pairs_u32(mp->d_sort_temp, mp->sort_temp_bytes, mp->d_keys, mp->d_skeys,
                                   mp->d_vals, mp->d_svals, n, bits, s, m->lc));
    }
    mark(FDEM_STAGE_ESTIMATE);
    if (tile) {
      if (mp->tb.job_counter == CNT_HEAVY) launch_tile_estimate_light(ep, mp->tb, mp->d_counters, st_out, s, m->lc);
      launch_tile_estimate(ep, mp->tb, mp->d_counters, st_out, L.pub, s, m->lc);
    }
    else launch_segreduce_estimate(ep, mp->d_counters, mp->d_counters, s, m->lc);
    mark(FDEM_STAGE_RAYCAST);
    if (raycast && !rc_beside && launch_status == FDEM_OK) launch_status = raycast_branch();
    if (raycast && launch_status == FDEM_OK) {
      // join: the map is only written here, after both the estimator and the DDA are done
      if (rc_beside) FDEM_CUDA_TRY(cudaStreamWaitEvent(s, mp->ev_rays, 0));
      launch_raycast_resolve(rcp, st_out, L.lt, mp->d_counters, m->cells, s, m->lc);
    }
    mark(FDEM_STAGE_COUNT);
    // counters + committed state -> host (mapped pinned memory), state made current, counters re-armed
    if (!L.pub.enabled)
      launch_publish(mp->d_counters, m->d_state, m->d_state + 1, host_slot, s, m->lc);
  }
  {
    cudaError_t le = cudaGetLastError();
    if (launch_status == FDEM_OK && le != cudaSuccess)
      launch_status = set_error(FDEM_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(le));
    if (launch_status != FDEM_OK) {
      mp->tile_dirty = true;
      mp->counters_dirty = true;
      return launch_status;
    }
  }
  FDEM_CUDA_TRY(cudaEventRecord(m->ev_scan[ticket % kResultRing], s));
  m->seq = ticket + 1;
  mp->last_ticket = ticket;
  m->geom_stale = true;
  mp->last_n = n;
  mp->last_raw = in.raw != nullptr;
  mp->last_had_work = true;
  mp->pending = true;
  return FDEM_OK;
}

// ── CUDA graph of one scan (tile path): explicit kernel nodes, patched in place ─────────
struct NodeArgs {
  KernelDesc d;
  void* args[12];
};

void scan_node_args(ScanLaunch& L, uint32_t n, NodeArgs out[GN_COUNT]) {
  out[GN_K1].d = desc_preprocess_bin(n);
  out[GN_K1].args[0] = &L.pp; out[GN_K1].args[1] = &L.st_cur; out[GN_K1].args[2] = &L.counters;
  out[GN_K1].args[3] = &L.pm; out[GN_K1].args[4] = &L.keys; out[GN_K1].args[5] = &L.vals;
  out[GN_K2].d = desc_commit();
  out[GN_K2].args[0] = &L.cp; out[GN_K2].args[1] = &L.st_cur; out[GN_K2].args[2] = &L.st_next;
  out[GN_K2].args[3] = &L.counters; out[GN_K2].args[4] = &L.lt;
  out[GN_SCATTER].d = desc_scatter_records(n);
  out[GN_SCATTER].args[0] = &L.sp;
  out[GN_K3L].d = desc_tile_estimate_light();
  out[GN_K3L].args[0] = &L.ep; out[GN_K3L].args[1] = &L.tb; out[GN_K3L].args[2] = &L.counters;
  out[GN_K3L].args[3] = &L.st_next;
  out[GN_K3T].d = desc_tile_estimate(L.tb.n_buckets, L.tb.bucket_bits);
  out[GN_K3T].args[0] = &L.ep; out[GN_K3T].args[1] = &L.tb; out[GN_K3T].args[2] = &L.counters;
  out[GN_K3T].args[3] = &L.st_next; out[GN_K3T].args[4] = &L.pub;
}

cudaKernelNodeParams node_params(NodeArgs& a) {
  cudaKernelNodeParams kp{};
  kp.func = const_cast<void*>(a.d.func);
  kp.gridDim = a.d.grid;
  kp.blockDim = a.d.block;
  kp.sharedMemBytes = static_cast<unsigned int>(a.d.smem);
  kp.kernelParams = a.args;
  kp.extra = nullptr;
  return kp;
}

void destroy_scan_graph(fdem_mapper* mp) {
  if (mp->sg.exec) cudaGraphExecDestroy(mp->sg.exec);
  if (mp->sg.graph) cudaGraphDestroy(mp->sg.graph);
  mp->sg = ScanGraph{};
}

fdem_status launch_scan_graph(fdem_mapper* mp, cudaStream_t s) {
  ScanLaunch& L = mp->launch;
  ScanGraph& G = mp->sg;
  NodeArgs na[GN_COUNT];
  scan_node_args(L, L.pp.n, na);
  const bool light = L.tb.job_counter == CNT_HEAVY;
  int chain[GN_COUNT], n_chain = 0;
  for (int i = 0; i < GN_COUNT; ++i)
    if (i != GN_K3L || light) chain[n_chain++] = i;
  if (!G.valid) {
    destroy_scan_graph(mp);
    FDEM_CUDA_TRY(cudaGraphCreate(&G.graph, 0));
    for (int c = 0; c < n_chain; ++c) {
      cudaKernelNodeParams kp = node_params(na[chain[c]]);
      FDEM_CUDA_TRY(cudaGraphAddKernelNode(&G.node[chain[c]], G.graph, nullptr, 0, &kp));
    }
    for (int c = 1; c < n_chain; ++c) {
      // programmatic edges: a node may launch as soon as its predecessor has called
      // griddepcontrol.launch_dependents; it blocks at griddepcontrol.wait until that one is done
      cudaGraphEdgeData ed{};
      if (mp->use_pdl) {
        ed.from_port = cudaGraphKernelNodePortProgrammatic;
        ed.type = cudaGraphDependencyTypeProgrammatic;
      }
      FDEM_CUDA_TRY(cudaGraphAddDependencies_v2(G.graph, &G.node[chain[c - 1]], &G.node[chain[c]], &ed, 1));
    }
    FDEM_CUDA_TRY(cudaGraphInstantiate(&G.exec, G.graph, 0));
    G.cached = L;
    G.valid = true;
  } else {
    const ScanLaunch& C = G.cached;
    const bool shared_changed = L.st_cur != C.st_cur || L.st_next != C.st_next ||
                                L.counters != C.counters || L.pm != C.pm || L.keys != C.keys ||
                                L.vals != C.vals || L.host_out != C.host_out ||
                                std::memcmp(&L.tb, &C.tb, sizeof(TileBuffers)) != 0;
    bool dirty[GN_COUNT];
    dirty[GN_K1] = shared_changed || std::memcmp(&L.pp, &C.pp, sizeof(L.pp)) != 0;
    dirty[GN_K2] = shared_changed || std::memcmp(&L.cp, &C.cp, sizeof(L.cp)) != 0 ||
                   std::memcmp(&L.lt, &C.lt, sizeof(L.lt)) != 0;
    dirty[GN_SCATTER] = shared_changed || std::memcmp(&L.sp, &C.sp, sizeof(L.sp)) != 0;
    dirty[GN_K3L] = shared_changed || std::memcmp(&L.ep, &C.ep, sizeof(L.ep)) != 0;
    dirty[GN_K3T] = dirty[GN_K3L] || std::memcmp(&L.pub, &C.pub, sizeof(L.pub)) != 0;
    for (int c = 0; c < n_chain; ++c) {
      const int i = chain[c];
      if (!dirty[i]) continue;
      cudaKernelNodeParams kp = node_params(na[i]);
      FDEM_CUDA_TRY(cudaGraphExecKernelNodeSetParams(G.exec, G.node[i], &kp));
    }
    G.cached = L;
  }
  FDEM_CUDA_TRY(cudaGraphLaunch(G.exec, s));
  return FDEM_OK;
}

// ── batched integration: S scans in ONE graph whose DAG lets scan k+1's front half (K1,
// commit, scatter — they touch only scan scratch and the geometry chain) run beside scan k's
// back half (back prologue + K3t — everything that writes the map).  Edges:
//   K1_i -> K2_i -> SC_i -> BP_i -> K3_i      (one scan)
//   K2_i -> K1_{i+1}                           (geometry after scan i's move)
//   K3_i -> BP_{i+1}                           (map layers, touched list)
//   K3_i -> K1_{i+2}                           (scratch set i & 1 is free again)
constexpr int kMaxBatch = 16;  // scans per batch graph
// BN_K3L: the warp-per-bucket pass, only in graphs built for the light scan shape
enum { BN_K1 = 0, BN_K2, BN_SC, BN_BP, BN_K3L, BN_K3, BN_COUNT };

struct BatchScan {
  ScanLaunch L;
  BackParams bp;
  uint32_t n;
};

}  // namespace

struct BatchGraphState {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t node[kMaxBatch][BN_COUNT] = {};
  cudaGraphNode_t zero = nullptr;   // head of the graph: all counter blocks -> 0
  cudaGraphNode_t report = nullptr; // counter blocks of scans 0..S-2 -> pinned host memory
  int S = 0;
  uint32_t tile_shape_key = 0;  // bucket bits the graph was built for
  BatchScan scan[kMaxBatch];    // this batch's kernel arguments
  BatchScan cached[kMaxBatch];  // what the executable graph currently holds
};

namespace {

void destroy_batch_graph(fdem_mapper* mp) {
  if (!mp->bg) return;
  for (int k = 0; k < 2; ++k) {
    if (mp->bg[k].exec) cudaGraphExecDestroy(mp->bg[k].exec);
    if (mp->bg[k].graph) cudaGraphDestroy(mp->bg[k].graph);
  }
  delete[] mp->bg;
  mp->bg = nullptr;
}

fdem_status ensure_batch_scratch(fdem_mapper* mp) {
  fdem_map* m = mp->map;
  if (!mp->d_ring) {
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_ring, (kMaxBatch + 1) * sizeof(DeviceState)));
    FDEM_CUDA_TRY(cudaMemset(mp->d_ring, 0, (kMaxBatch + 1) * sizeof(DeviceState)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_move, 2 * sizeof(MoveRecord)));
    FDEM_CUDA_TRY(cudaMemset(mp->d_move, 0, 2 * sizeof(MoveRecord)));
    FDEM_CUDA_TRY(cudaHostAlloc(&mp->h_batch_counters, kMaxBatch * CNT_COUNT * sizeof(uint32_t), cudaHostAllocDefault));
    std::memset(mp->h_batch_counters, 0, kMaxBatch * CNT_COUNT * sizeof(uint32_t));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_batch_counters, kMaxBatch * CNT_COUNT * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMemset(mp->d_batch_counters, 0, kMaxBatch * CNT_COUNT * sizeof(uint32_t)));
    const size_t nb = std::max<size_t>(mp->max_buckets, 1) * sizeof(uint32_t);
    cudaFree(mp->tb2.bucket_count); cudaFree(mp->tb2.bucket_offset); cudaFree(mp->tb2.bucket_cursor);
    cudaFree(mp->tb2.bucket_list); cudaFree(mp->tb2.heavy_list);
    FDEM_CUDA_TRY(cudaMalloc(&mp->tb2.heavy_list, nb * 4));
    FDEM_CUDA_TRY(cudaMalloc(&mp->tb2.bucket_count, nb));
    FDEM_CUDA_TRY(cudaMalloc(&mp->tb2.bucket_offset, nb));
    FDEM_CUDA_TRY(cudaMalloc(&mp->tb2.bucket_cursor, nb));
    FDEM_CUDA_TRY(cudaMalloc(&mp->tb2.bucket_list, nb * 4));
    FDEM_CUDA_TRY(cudaMemset(mp->tb2.bucket_count, 0, nb));
    FDEM_CUDA_TRY(cudaMemset(mp->tb2.bucket_cursor, 0, nb));
  }
  if (mp->cap2 < mp->cap) {
    FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
    cudaFree(mp->d_pm2);
    cudaFree(mp->d_keys2);
    cudaFree(mp->tb2.records);
    mp->d_pm2 = nullptr; mp->d_keys2 = nullptr; mp->tb2.records = nullptr;
    mp->cap2 = 0;
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_pm2, mp->cap * sizeof(float4)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->d_keys2, mp->cap * sizeof(uint32_t)));
    FDEM_CUDA_TRY(cudaMalloc(&mp->tb2.records, mp->cap * sizeof(CellRecord)));
    mp->cap2 = mp->cap;
  }
  return FDEM_OK;
}

// point scan i's kernel arguments (built against the primary scratch set) at scratch set
// i & 1, its slots of the state ring, and the deferred-commit / back-prologue split
void retarget_for_batch(fdem_mapper* mp, int i, BatchScan& b) {
  fdem_map* m = mp->map;
  ScanLaunch& L = b.L;
  const int q = i & 1;
  TileBuffers tb = q ? mp->tb2 : mp->tb;
  tb.n_buckets = mp->tb.n_buckets;
  tb.bucket_bits = mp->tb.bucket_bits;
  tb.job_counter = mp->tb.job_counter;
  float4* pm = q ? mp->d_pm2 : mp->d_pm;
  uint32_t* keys = q ? mp->d_keys2 : mp->d_keys;
  // one counter block per scan of the batch, all zeroed by ONE memset node at the head of the
  // graph: no scan has to re-arm its counters, so only the last scan needs the last-CTA ticket
  uint32_t* counters = mp->d_batch_counters + static_cast<size_t>(i) * CNT_COUNT;
  const DeviceState* st_in = i == 0 ? m->d_state : mp->d_ring + i;
  DeviceState* st_out = mp->d_ring + i + 1;
  L.pp.bucket_count = tb.bucket_count;
  L.pm = pm; L.keys = keys; L.counters = counters;
  L.st_cur = st_in;            // K1 / K2 read the geometry the previous scan's commit left
  L.st_next = st_out;
  L.cp.tb = tb;
  L.cp.defer = 1;
  L.cp.move_out = mp->d_move + q;
  L.cp.obstacle = nullptr;
  L.sp.keys = keys; L.sp.pm = pm; L.sp.tb = tb; L.sp.counters = counters;
  L.sp.obstacle = nullptr;     // the back prologue resets the obstacle cells
  L.ep.pm = pm;
  L.tb = tb;
  L.pub.st_cur = m->d_state;   // the batch's last scan commits the state and reports to the host
  b.bp = BackParams{};
  b.bp.counters = counters;
  b.bp.st_cur = st_in;         // the previous scan's slot: its estimator left the touched count there
  b.bp.st_out = st_out;
  b.bp.move = mp->d_move + q;
  b.bp.obstacle = L.ep.L.obstacle;
  b.bp.touched_keys = m->d_touched_keys;
  b.bp.invalid_key = L.pp.invalid_key;
  b.bp.clear_policy = L.cp.clear_policy;
  b.bp.obstacle_cells = m->cells;
}

void batch_node_args(BatchScan& b, NodeArgs out[BN_COUNT]) {
  ScanLaunch& L = b.L;
  out[BN_K1].d = desc_preprocess_bin(b.n);
  out[BN_K1].args[0] = &L.pp; out[BN_K1].args[1] = &L.st_cur; out[BN_K1].args[2] = &L.counters;
  out[BN_K1].args[3] = &L.pm; out[BN_K1].args[4] = &L.keys; out[BN_K1].args[5] = &L.vals;
  out[BN_K2].d = desc_commit();
  out[BN_K2].args[0] = &L.cp; out[BN_K2].args[1] = &L.st_cur; out[BN_K2].args[2] = &L.st_next;
  out[BN_K2].args[3] = &L.counters; out[BN_K2].args[4] = &L.lt;
  out[BN_SC].d = desc_scatter_records(b.n);
  out[BN_SC].args[0] = &L.sp;
  out[BN_BP].d = desc_back_prologue();
  out[BN_BP].args[0] = &b.bp; out[BN_BP].args[1] = &L.lt;
  out[BN_K3L].d = desc_tile_estimate_light();
  out[BN_K3L].args[0] = &L.ep; out[BN_K3L].args[1] = &L.tb; out[BN_K3L].args[2] = &L.counters;
  out[BN_K3L].args[3] = &L.st_next;
  out[BN_K3].d = desc_tile_estimate(L.tb.n_buckets, L.tb.bucket_bits);
  out[BN_K3].args[0] = &L.ep; out[BN_K3].args[1] = &L.tb; out[BN_K3].args[2] = &L.counters;
  out[BN_K3].args[3] = &L.st_next; out[BN_K3].args[4] = &L.pub;
}

fdem_status launch_batch_graph(fdem_mapper* mp, int S, cudaStream_t s) {
  BatchGraphState& G = mp->bg[mp->bg_flip];
  const bool light = mp->tb.job_counter == CNT_HEAVY;
  const uint32_t shape_key = mp->tb.bucket_bits | (light ? 0x100u : 0u);
  const bool rebuild = !G.exec || G.S != S || G.tile_shape_key != shape_key;
  int chain[BN_COUNT], n_chain = 0;
  for (int k = 0; k < BN_COUNT; ++k)
    if (k != BN_K3L || light) chain[n_chain++] = k;
  if (rebuild) {
    if (G.exec) cudaGraphExecDestroy(G.exec);
    if (G.graph) cudaGraphDestroy(G.graph);
    G.exec = nullptr; G.graph = nullptr;
    FDEM_CUDA_TRY(cudaGraphCreate(&G.graph, 0));
    {
      cudaMemsetParams zp{};
      zp.dst = mp->d_batch_counters;
      zp.value = 0;
      zp.elementSize = 4;
      zp.width = static_cast<size_t>(kMaxBatch) * CNT_COUNT;
      zp.height = 1;
      FDEM_CUDA_TRY(cudaGraphAddMemsetNode(&G.zero, G.graph, nullptr, 0, &zp));
    }
    for (int i = 0; i < S; ++i) {
      NodeArgs na[BN_COUNT];
      batch_node_args(G.scan[i], na);
      for (int c = 0; c < n_chain; ++c) {
        cudaKernelNodeParams kp = node_params(na[chain[c]]);
        FDEM_CUDA_TRY(cudaGraphAddKernelNode(&G.node[i][chain[c]], G.graph, nullptr, 0, &kp));
      }
    }
    FDEM_CUDA_TRY(cudaGraphAddDependencies(G.graph, &G.zero, &G.node[0][BN_K1], 1));
    auto edge = [&](cudaGraphNode_t a, cudaGraphNode_t b) -> cudaError_t {
      return cudaGraphAddDependencies(G.graph, &a, &b, 1);
    };
    // programmatic edges along the back chain (BP_i -> K3_i -> BP_{i+1}): the successor is
    // launched early and blocks at griddepcontrol.wait, so its launch latency and prologue
    // hide behind the predecessor (FDEM_BATCH_PDL=1)
    static const bool pdl = [] { const char* e = std::getenv("FDEM_BATCH_PDL"); return e && e[0] == '1'; }();
    auto pedge = [&](cudaGraphNode_t a, cudaGraphNode_t b) -> cudaError_t {
      if (!pdl) return cudaGraphAddDependencies(G.graph, &a, &b, 1);
      cudaGraphEdgeData ed{};
      ed.from_port = cudaGraphKernelNodePortProgrammatic;
      ed.type = cudaGraphDependencyTypeProgrammatic;
      return cudaGraphAddDependencies_v2(G.graph, &a, &b, &ed, 1);
    };
    for (int i = 0; i < S; ++i) {
      for (int c = 1; c < n_chain; ++c) {
        const int k = chain[c];
        if (k == BN_K3 || k == BN_K3L) FDEM_CUDA_TRY(pedge(G.node[i][chain[c - 1]], G.node[i][k]));
        else FDEM_CUDA_TRY(edge(G.node[i][chain[c - 1]], G.node[i][k]));
      }
      if (i + 1 < S) {
        FDEM_CUDA_TRY(edge(G.node[i][BN_K2], G.node[i + 1][BN_K1]));
        FDEM_CUDA_TRY(pedge(G.node[i][BN_K3], G.node[i + 1][BN_BP]));
      }
      if (i + 2 < S) FDEM_CUDA_TRY(edge(G.node[i][BN_K3], G.node[i + 2][BN_K1]));
    }
    if (S > 1) {
      // every scan's statistics reach the host, like in the per-scan pipeline: the last scan's
      // through its publish, the others' through one copy of their counter blocks, issued as
      // soon as the second-to-last scan is done (beside the last scan, off the critical path)
      FDEM_CUDA_TRY(cudaGraphAddMemcpyNode1D(&G.report, G.graph, &G.node[S - 2][BN_K3], 1, mp->h_batch_counters,
                                             mp->d_batch_counters, sizeof(uint32_t) * CNT_COUNT * (S - 1),
                                             cudaMemcpyDeviceToHost));
    }
    FDEM_CUDA_TRY(cudaGraphInstantiate(&G.exec, G.graph, 0));
    G.S = S;
    G.tile_shape_key = shape_key;
  } else {
    // patch only the nodes whose arguments differ from what the executable graph holds
    for (int i = 0; i < S; ++i) {
      const BatchScan& c = G.cached[i];
      const BatchScan& b = G.scan[i];
      const ScanLaunch& L = b.L;
      const ScanLaunch& C = c.L;
      const bool shared = L.st_cur != C.st_cur || L.st_next != C.st_next || L.counters != C.counters ||
                          L.pm != C.pm || L.keys != C.keys || L.vals != C.vals || b.n != c.n ||
                          std::memcmp(&L.tb, &C.tb, sizeof(TileBuffers)) != 0;
      bool dirty[BN_COUNT];
      dirty[BN_K1] = shared || std::memcmp(&L.pp, &C.pp, sizeof(L.pp)) != 0;
      dirty[BN_K2] = shared || std::memcmp(&L.cp, &C.cp, sizeof(L.cp)) != 0 ||
                     std::memcmp(&L.lt, &C.lt, sizeof(L.lt)) != 0;
      dirty[BN_SC] = shared || std::memcmp(&L.sp, &C.sp, sizeof(L.sp)) != 0;
      dirty[BN_BP] = shared || std::memcmp(&b.bp, &c.bp, sizeof(b.bp)) != 0 ||
                     std::memcmp(&L.lt, &C.lt, sizeof(L.lt)) != 0;
      dirty[BN_K3L] = shared || std::memcmp(&L.ep, &C.ep, sizeof(L.ep)) != 0;
      dirty[BN_K3] = dirty[BN_K3L] || std::memcmp(&L.pub, &C.pub, sizeof(L.pub)) != 0;
      NodeArgs na[BN_COUNT];
      batch_node_args(G.scan[i], na);
      for (int c = 0; c < n_chain; ++c) {
        const int k = chain[c];
        if (!dirty[k]) continue;
        cudaKernelNodeParams kp = node_params(na[k]);
        FDEM_CUDA_TRY(cudaGraphExecKernelNodeSetParams(G.exec, G.node[i][k], &kp));
      }
    }
  }
  for (int i = 0; i < S; ++i) G.cached[i] = G.scan[i];
  FDEM_CUDA_TRY(cudaGraphLaunch(G.exec, s));
  return FDEM_OK;
}

// Bucket shape for the NEXT scans from a finished scan's statistics.  K3t runs one CTA per
// non-empty bucket: when a scan fills only a few dozen 1024-cell buckets (a dense cloud on a
// small patch: an RGB-D frame, a coarse map) most SMs idle while a few CTAs grind through
// thousands of records, so the mapper switches to 256-cell buckets with one thread per cell
// (4x the CTAs, 4x the threads per cell); when 256-cell buckets would mean thousands of jobs
// it switches back.  The L1 scratch is all-zero between scans (K3t re-arms what it consumed)
// and sized for the smallest shape, so switching needs no clearing; in-flight scans keep the
// shape they were enqueued with.
//
// The other extreme — a scan spread thin over a large map (a LiDAR sweep on a 0.05 m global map:
// thousands of buckets holding a few dozen cells each) — takes the "light" shape: a first pass
// handles every bucket with <= 128 records with one WARP (tile_estimate_light_kernel), and the
// CTA-per-bucket kernel only sees what that pass left over.
void adapt_bucket_shape(fdem_mapper* mp, uint32_t nonempty_buckets, uint32_t n_cells) {
  if (!mp->use_tile || nonempty_buckets == 0) return;
  uint32_t bits = mp->tb.bucket_bits;
  if (mp->bucket_bits_auto) {
    if (bits == 10u && nonempty_buckets < 160u) bits = 8u;
    else if (bits == 8u && nonempty_buckets > 800u) bits = 10u;
  }
  bool light = mp->tb.job_counter == CNT_HEAVY;
  if (mp->tile_light_auto) {
    const uint32_t cells_per_bucket = n_cells / nonempty_buckets;
    if (!light && bits == 10u && nonempty_buckets >= 600u && cells_per_bucket <= 64u) light = true;
    else if (light && (nonempty_buckets < 400u || cells_per_bucket > 96u)) light = false;
  } else {
    light = mp->tile_light_pin;
  }
  if (bits != 10u) light = false;
  if (bits == mp->tb.bucket_bits && light == (mp->tb.job_counter == CNT_HEAVY)) return;
  mp->tb.bucket_bits = bits;
  mp->tb.n_buckets = static_cast<uint32_t>((mp->map->cells + (1ull << bits) - 1) >> bits);
  mp->tb.job_counter = light ? CNT_HEAVY : CNT_BUCKETS;
  mp->sg.valid = false;  // K3t's function, grid and chain change: rebuild the scan graph
}

fdem_status finish_scan(fdem_mapper* mp, fdem_scan_stats* stats) {
  fdem_map* m = mp->map;
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (!mp->ev_scans.empty()) drain_stage_events(mp);
  if (mp->last_had_work) {
    const ScanResult& r = m->h_result[mp->last_ticket % kResultRing];
    m->geom = r.state.geom;
    m->geom_stale = false;
    apply_state_flags(m, r.state.flags);
    mp->last.n_input = mp->last_raw ? r.counters[CNT_FINITE] : mp->last_n;
    mp->last.n_kept = r.counters[CNT_KEPT];
    mp->last.n_cells = r.counters[CNT_CELLS];
    mp->last.n_voxels = r.counters[CNT_VOXELS];
    mp->last.integrated = r.counters[CNT_KEPT] > 0 ? 1 : 0;
    mp->last.voxel_box_violations = static_cast<int32_t>(r.counters[CNT_VOX_VIOLATION]);
    adapt_bucket_shape(mp, r.counters[CNT_BUCKETS], r.counters[CNT_CELLS]);
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

fdem_status fdem_config_validate(fdem_config* c, int32_t* n_clamped) {
  FDEM_REQUIRE(c != nullptr, "config is null");
  int32_t n = 0;
  // fatal (config_fastdem.cpp:131-137)
  if (c->kalman_min_variance >= c->kalman_max_variance)
    return set_error(FDEM_ERR_INVALID_ARGUMENT, "mapping.kalman: min_variance >= max_variance");
  auto fix = [&](bool bad, float& v, float to) { if (bad) { v = to; ++n; } };
  if (c->raycasting_enabled) {  // :148-179
    fix(c->rc_height_conflict_threshold <= 0.0f, c->rc_height_conflict_threshold, 0.05f);
    fix(c->rc_log_odds_observed <= 0.0f, c->rc_log_odds_observed, 0.4f);
    fix(c->rc_log_odds_ghost <= 0.0f, c->rc_log_odds_ghost, 0.2f);
    fix(c->rc_log_odds_max <= 0.0f, c->rc_log_odds_max, 2.0f);
    fix(c->rc_clear_threshold >= 0.0f, c->rc_clear_threshold, -1.0f);
  }
  fix(c->kalman_min_variance <= 0.0f, c->kalman_min_variance, 0.0001f);   // :181-187
  fix(c->kalman_process_noise < 0.0f, c->kalman_process_noise, 0.0f);    // :188-194
  if (c->p2_elevation_marker < 0 || c->p2_elevation_marker > 4) {        // :195-196
    c->p2_elevation_marker = std::min(std::max(c->p2_elevation_marker, 0), 4);
    ++n;
  }
  for (int i = 0; i < 5; ++i)                                            // :199-207
    if (c->p2_dn[i] < 0.0f || c->p2_dn[i] > 1.0f) {
      c->p2_dn[i] = std::min(std::max(c->p2_dn[i], 0.0f), 1.0f);
      ++n;
    }
  for (int i = 1; i < 5; ++i)                                            // :208-216
    if (c->p2_dn[i - 1] > c->p2_dn[i])
      return set_error(FDEM_ERR_INVALID_ARGUMENT, "mapping.p2: markers must be sorted (dn0 <= ... <= dn4)");
  fix(c->lidar_range_noise <= 0.0f, c->lidar_range_noise, 0.02f);        // :219-224
  fix(c->lidar_angular_noise < 0.0f, c->lidar_angular_noise, 0.0f);      // :225-230
  fix(c->constant_uncertainty <= 0.0f, c->constant_uncertainty, 0.1f);   // :231-237
  fix(c->rgbd_normal_a < 0.0f, c->rgbd_normal_a, 0.0f);                  // :238-257
  fix(c->rgbd_normal_b < 0.0f, c->rgbd_normal_b, 0.0f);
  fix(c->rgbd_normal_c < 0.0f, c->rgbd_normal_c, 0.0f);
  fix(c->rgbd_lateral_factor < 0.0f, c->rgbd_lateral_factor, 0.0f);
  if (n_clamped) *n_clamped = n;
  return FDEM_OK;
}

// ───────────────────────────── map ───────────────────────────────────────────

// ElevationMap::setGeometry(float, float, float) widens to double, then nanoGrid setGeometry:
// size = round(length / res), length = size * res, position = 0, start index = 0
// (elevation_map.hpp:112-116)
static bool make_geometry(float width, float height, float resolution, int32_t row_begin,
                          int32_t row_end, GridGeom* out) {
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
  if (g.rows <= 0 || g.cols <= 0 || row_begin < 0 || row_end > g.rows || row_begin >= row_end) return false;
  g.row_begin = row_begin;
  g.row_end = row_end;
  *out = g;
  return true;
}

static fdem_status map_init(fdem_map* m, void* stream) {
  if (stream) {
    m->stream = static_cast<cudaStream_t>(stream);
  } else {
    // the map's own stream carries everything that writes the map; it gets the highest priority so
    // that, when a mapper overlaps the next scan's front half on its side stream, the CTAs of the
    // map-writing kernels are placed first
    int prio_least = 0, prio_greatest = 0;
    FDEM_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    FDEM_CUDA_TRY(cudaStreamCreateWithPriority(&m->stream, cudaStreamNonBlocking, prio_greatest));
    m->own_stream = true;
  }
  FDEM_CUDA_TRY(cudaMalloc(&m->d_state, 2 * sizeof(DeviceState)));
  FDEM_CUDA_TRY(cudaMemset(m->d_state, 0, 2 * sizeof(DeviceState)));
  FDEM_CUDA_TRY(cudaMalloc(&m->d_flag, 4 * sizeof(uint32_t)));
  FDEM_CUDA_TRY(cudaHostAlloc(&m->h_result, kResultRing * sizeof(ScanResult), cudaHostAllocMapped));
  std::memset(m->h_result, 0, kResultRing * sizeof(ScanResult));
  FDEM_CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&m->d_result_host), m->h_result, 0));
  for (int i = 0; i < kResultRing; ++i)
    FDEM_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_scan[i], cudaEventDisableTiming));
  FDEM_TRY(push_state(m, 0));
  // ElevationMap() basic layers (elevation_map.hpp:99-103), then clearAll()
  for (const char* name : {"elevation", "elevation_min", "elevation_max"})
    FDEM_TRY(add_layer(m, name, kNaN));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  return FDEM_OK;
}

fdem_status fdem_map_create_stripe(float width, float height, float resolution, int32_t row_begin,
                                   int32_t row_end, int32_t device, void* stream, fdem_map** out) {
  FDEM_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  FDEM_REQUIRE(resolution > 0.0f && width > 0.0f && height > 0.0f, "bad geometry");
  int ndev = 0;
  FDEM_CUDA_TRY(cudaGetDeviceCount(&ndev));
  FDEM_REQUIRE(device >= 0 && device < ndev, "bad device ordinal");
  GridGeom g{};
  if (!make_geometry(width, height, resolution, row_begin, row_end, &g))
    return set_error(FDEM_ERR_INVALID_ARGUMENT, "bad size or row stripe");
  DeviceGuard dg(device);  // the caller's current device is left as it was
  FDEM_REQUIRE(dg.ok, "cannot select the device");
  fdem_map* m = new (std::nothrow) fdem_map();
  if (!m) return set_error(FDEM_ERR_OUT_OF_MEMORY, "host allocation failed");
  m->device = device;
  m->geom = g;
  m->cells = static_cast<size_t>(g.row_end - g.row_begin) * g.cols;
  const fdem_status st = map_init(m, stream);
  if (st != FDEM_OK) {
    const std::string why = g_last_error;
    fdem_map_destroy(m);  // tolerates the members that were never created
    return set_error(st, why);
  }
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
  // a FastDEM still bound to this map must not be left with a dangling pointer: it is detached
  // (its calls then fail with FDEM_ERR_INVALID_ARGUMENT; fdem_mapper_destroy stays valid)
  for (fdem_mapper* mp : m->mappers) {
    cudaStreamSynchronize(m->stream);
    mp->map = nullptr;
  }
  m->mappers.clear();
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (auto& l : m->layers) cudaFree(l.d);
  cudaFree(m->d_state);
  cudaFree(m->d_flag);
  cudaFree(m->d_touched_keys);
  cudaFree(m->d_touched_minz);
  cudaFree(m->d_ray_min_enc);
  cudaFree(m->d_hits);
  cudaFree(m->d_pack);
  cudaFree(m->d_pack_cols);
  cudaFree(m->d_tmp[0]);
  cudaFree(m->d_tmp[1]);
  cudaFreeHost(m->h_result);
  for (cudaEvent_t e : m->ev_scan)
    if (e) cudaEventDestroy(e);
  if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
  (void)cudaGetLastError();
  delete m;
  return FDEM_OK;
}

static fdem_status mapper_resize_for_map(fdem_mapper* mp);

// GridMap::setGeometry + clearAll (elevation_map.hpp:112-116) IN PLACE: same handle, every
// existing layer is re-allocated for the new size and reset to NaN, position / start index
// return to 0.  FastDEM objects bound to the map stay valid (the reference's FastDEM holds an
// ElevationMap& and io::loadNpz calls setGeometry on it, io_npz.cpp:440-): their map-sized
// scratch is rebuilt here.
fdem_status fdem_map_set_geometry(fdem_map* m, float width, float height, float resolution) {
  FDEM_REQUIRE(m, "null map");
  FDEM_REQUIRE(resolution > 0.0f && width > 0.0f && height > 0.0f, "bad geometry");
  DeviceGuard dg(m->device);
  FDEM_REQUIRE(m->geom.row_begin == 0 && m->geom.row_end == m->geom.rows,
               "setGeometry is not valid on a row stripe");
  GridGeom g{};
  if (!make_geometry(width, height, resolution, -1, -1, &g))
    return set_error(FDEM_ERR_INVALID_ARGUMENT, "bad size");
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  for (fdem_mapper* mp : m->mappers) {
    if (mp->copy_stream) FDEM_CUDA_TRY(cudaStreamSynchronize(mp->copy_stream));
    if (mp->aux_stream) FDEM_CUDA_TRY(cudaStreamSynchronize(mp->aux_stream));
  }
  const size_t cells = static_cast<size_t>(g.rows) * g.cols;
  if (cells != m->cells) {
    for (auto& l : m->layers) {
      cudaFree(l.d);
      l.d = nullptr;
      FDEM_CUDA_TRY(cudaMalloc(&l.d, std::max<size_t>(cells, 1) * sizeof(float)));
    }
    // map-sized scratch is re-created on demand
    cudaFree(m->d_ray_min_enc); m->d_ray_min_enc = nullptr;
    cudaFree(m->d_hits); m->d_hits = nullptr;
    cudaFree(m->d_tmp[0]); cudaFree(m->d_tmp[1]); m->d_tmp[0] = m->d_tmp[1] = nullptr;
    cudaFree(m->d_pack_cols); m->d_pack_cols = nullptr;
    ++m->layer_epoch;
  } else if (m->d_ray_min_enc) {
    launch_fill_u32(m->d_ray_min_enc, cells, 0u, m->stream, m->lc);
    launch_fill_u32(m->d_hits, cells, 0u, m->stream, m->lc);
  }
  m->cells = cells;
  m->geom = g;
  m->geom_stale = false;
  m->state_flags &= ~static_cast<uint32_t>(SF_OBSTACLE_DIRTY);
  m->obstacle_full_clear = false;
  for (auto& l : m->layers) launch_fill(l.d, m->cells, kNaN, m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  FDEM_TRY(push_state(m, 0));  // touched list empty: nothing of the old map survives
  FDEM_CUDA_TRY(cudaMemcpy(m->d_state + 1, m->d_state, sizeof(DeviceState), cudaMemcpyDeviceToDevice));
  for (fdem_mapper* mp : m->mappers) FDEM_TRY(mapper_resize_for_map(mp));
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
  FDEM_CUDA_TRY(cudaMemcpy(&st, m->d_state, sizeof(st), cudaMemcpyDeviceToHost));
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
  launch_move_only(m->d_state, m->d_state + 1, x, y, clear_policy, layer_table(m), m->d_flag,
                   m->stream, m->lc);
  FDEM_CUDA_TRY(cudaGetLastError());
  FDEM_CUDA_TRY(cudaMemcpyAsync(m->d_state, m->d_state + 1, sizeof(DeviceState),
                                cudaMemcpyDeviceToDevice, m->stream));
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
  DeviceGuard dg(m->device);
  *exists = find_visible_layer(m, name) ? 1 : 0;
  return FDEM_OK;
}

fdem_status fdem_map_layer_add(fdem_map* m, const char* name, float fill) {
  FDEM_REQUIRE(m && name && name[0], "null argument");
  DeviceGuard dg(m->device);
  if (Layer* l = find_layer(m, name)) l->hidden_until = 0;  // GridMap::add on an existing name refills it
  return add_layer(m, name, fill);
}

fdem_status fdem_map_layer_count(fdem_map* m, int32_t* count) {
  FDEM_REQUIRE(m && count, "null argument");
  DeviceGuard dg(m->device);
  settle_hidden_layers(m);
  int32_t n = 0;
  for (const auto& l : m->layers) n += l.hidden_until ? 0 : 1;
  *count = n;
  return FDEM_OK;
}

fdem_status fdem_map_layer_name(fdem_map* m, int32_t i, char* buf, int32_t cap) {
  FDEM_REQUIRE(m && buf && cap > 0, "null argument");
  DeviceGuard dg(m->device);
  settle_hidden_layers(m);
  int32_t k = 0;
  for (const auto& l : m->layers) {
    if (l.hidden_until) continue;
    if (k++ == i) {
      std::snprintf(buf, cap, "%s", l.name.c_str());
      return FDEM_OK;
    }
  }
  return set_error(FDEM_ERR_INVALID_ARGUMENT, "layer index out of range");
}

fdem_status fdem_map_layer_download(fdem_map* m, const char* name, float* dst) {
  FDEM_REQUIRE(m && name && dst, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_visible_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  FDEM_CUDA_TRY(cudaMemcpyAsync(dst, l->d, m->cells * sizeof(float), cudaMemcpyDefault, m->stream));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  return FDEM_OK;
}

fdem_status fdem_map_layer_upload(fdem_map* m, const char* name, const float* src) {
  FDEM_REQUIRE(m && name && src, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_visible_layer(m, name);
  if (!l) return set_error(FDEM_ERR_NO_LAYER, std::string("no such layer: ") + name);
  FDEM_CUDA_TRY(cudaMemcpyAsync(l->d, src, m->cells * sizeof(float), cudaMemcpyDefault, m->stream));
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (l->name == "obstacle") m->obstacle_full_clear = true;
  return FDEM_OK;
}

fdem_status fdem_map_layer_device_ptr(fdem_map* m, const char* name, float** dptr) {
  FDEM_REQUIRE(m && name && dptr, "null argument");
  DeviceGuard dg(m->device);
  Layer* l = find_visible_layer(m, name);
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
  Layer* l = find_visible_layer(m, name);
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
  Layer* l = find_visible_layer(m, name);
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
  Layer* l = find_visible_layer(m, name);
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

// (re)build everything in the mapper that is sized by the map: bucket tables of both scratch
// sets, graphs.  Called at creation and after fdem_map_set_geometry.
static fdem_status mapper_resize_for_map(fdem_mapper* mp) {
  fdem_map* map = mp->map;
  destroy_scan_graph(mp);
  destroy_batch_graph(mp);
  for (TileBuffers* tb : {&mp->tb, &mp->tb2}) {
    cudaFree(tb->bucket_count); cudaFree(tb->bucket_offset); cudaFree(tb->bucket_cursor); cudaFree(tb->bucket_list);
    cudaFree(tb->heavy_list);
    tb->bucket_count = tb->bucket_offset = tb->bucket_cursor = nullptr;
    tb->bucket_list = tb->heavy_list = nullptr;
  }
  // the second scratch set / state ring of batched integration is rebuilt on its next use
  cudaFree(mp->d_ring); mp->d_ring = nullptr;
  cudaFree(mp->d_move); mp->d_move = nullptr;
  cudaFree(mp->d_batch_counters); mp->d_batch_counters = nullptr;
  cudaFreeHost(mp->h_batch_counters); mp->h_batch_counters = nullptr;
  const uint32_t bits = mp->tb.bucket_bits;
  mp->max_buckets = static_cast<uint32_t>((map->cells + 255u) >> 8);
  mp->tb.n_buckets = static_cast<uint32_t>((map->cells + (1ull << bits) - 1) >> bits);
  const size_t nb = std::max<size_t>(mp->max_buckets, 1) * sizeof(uint32_t);
  FDEM_CUDA_TRY(cudaMalloc(&mp->tb.bucket_count, nb));
  FDEM_CUDA_TRY(cudaMalloc(&mp->tb.bucket_offset, nb));
  FDEM_CUDA_TRY(cudaMalloc(&mp->tb.bucket_cursor, nb));
  FDEM_CUDA_TRY(cudaMalloc(&mp->tb.bucket_list, nb * 4));
  FDEM_CUDA_TRY(cudaMalloc(&mp->tb.heavy_list, nb * 4));
  mp->tile_dirty = true;
  mp->counters_dirty = true;
  mp->cached_epoch = ~0ull;
  mp->last_had_work = false;
  // setGeometry + clearAll leaves the estimator layers in place, all NaN (they exist already)
  return ensure_estimator_layers(map, mp->cfg);
}

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
  mp->device = map->device;
  mp->cfg = c;
  fdem_status st = ensure_estimator_layers(map, c);
  if (st != FDEM_OK) { delete mp; return st; }
  cudaError_t e = cudaMalloc(&mp->d_counters, CNT_COUNT * sizeof(uint32_t));
  if (e != cudaSuccess) { delete mp; return set_error(FDEM_ERR_CUDA, cudaGetErrorString(e)); }
  {
    cudaError_t e3 = cudaStreamCreateWithFlags(&mp->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < kStageRing && e3 == cudaSuccess; ++i)
      e3 = cudaEventCreateWithFlags(&mp->ev_copied[i], cudaEventDisableTiming);
    if (e3 != cudaSuccess) {
      fdem_mapper_destroy(mp);
      return set_error(FDEM_ERR_CUDA, std::string("copy stream setup failed: ") + cudaGetErrorString(e3));
    }
  }
  {
    cudaError_t e4 = cudaStreamCreateWithFlags(&mp->aux_stream, cudaStreamNonBlocking);
    if (e4 == cudaSuccess) e4 = cudaEventCreateWithFlags(&mp->ev_geom, cudaEventDisableTiming);
    if (e4 == cudaSuccess) e4 = cudaEventCreateWithFlags(&mp->ev_rays, cudaEventDisableTiming);
    if (e4 != cudaSuccess) {
      fdem_mapper_destroy(mp);
      return set_error(FDEM_ERR_CUDA, std::string("aux stream setup failed: ") + cudaGetErrorString(e4));
    }
  }
  {
    // tile path scratch: one counter / offset / cursor / list slot per 1024-cell bucket
    const char* env = std::getenv("FDEM_CELL_SORT");
    mp->use_tile = !(env && std::string(env) == "cub");
    const char* genv = std::getenv("FDEM_GRAPH");
    mp->use_graph = !(genv && std::string(genv) == "0");
    const char* penv = std::getenv("FDEM_PDL");
    mp->use_pdl = penv && std::string(penv) == "1";  // measured: no gain inside a graph; opt-in
    const char* venv = std::getenv("FDEM_VOXEL_SORT");
    // measured on B200 (profiles/, r2): cub's onesweep sorts the 32-bit voxel keys of a 1M-point
    // scan in 72 us; our MSD sort is exact but its CTA-per-row level costs 150 us on the rows a
    // dense scan crowds.  The library sort stays the default; FDEM_VOXEL_SORT=msd selects ours.
    mp->voxel_sort_library = !(venv && std::string(venv) == "msd");
    // K3t bucket shape (kernels_tile.cu): 1024-cell buckets to start with; after every scan
    // whose statistics the host has seen the shape follows the scan's density (set_bucket_bits
    // below).  FDEM_BUCKET_BITS=8|9|10 pins it.
    const char* benv = std::getenv("FDEM_BUCKET_BITS");
    uint32_t bits = 10u;
    if (benv) {
      const int v = std::atoi(benv);
      if (v >= 8 && v <= 10) { bits = static_cast<uint32_t>(v); mp->bucket_bits_auto = false; }
    }
    mp->tb.bucket_bits = bits;
    // FDEM_TILE_LIGHT=0|1 pins the warp-per-bucket pass off / on (1024-cell buckets only)
    const char* lenv = std::getenv("FDEM_TILE_LIGHT");
    mp->tb.job_counter = CNT_BUCKETS;
    if (lenv && (lenv[0] == '0' || lenv[0] == '1')) {
      mp->tile_light_auto = false;
      mp->tile_light_pin = lenv[0] == '1';
      if (mp->tile_light_pin && bits == 10u) mp->tb.job_counter = CNT_HEAVY;
    }
    mp->device = map->device;
    cudaError_t e2 = static_cast<cudaError_t>(tile_estimate_configure());
    if (e2 != cudaSuccess) {
      fdem_mapper_destroy(mp);
      return set_error(FDEM_ERR_CUDA, std::string("tile path setup failed: ") + cudaGetErrorString(e2));
    }
    map->mappers.push_back(mp);
    const fdem_status rs = mapper_resize_for_map(mp);
    if (rs != FDEM_OK) {
      const std::string why = g_last_error;
      fdem_mapper_destroy(mp);
      return set_error(rs, why);
    }
  }
  *out = mp;
  return FDEM_OK;
}

fdem_status fdem_mapper_destroy(fdem_mapper* mp) {
  if (!mp) return FDEM_OK;
  DeviceGuard dg(mp->device);
  if (mp->map) {
    cudaStreamSynchronize(mp->map->stream);
    auto& v = mp->map->mappers;
    v.erase(std::remove(v.begin(), v.end(), mp), v.end());
  }
  if (mp->copy_stream) cudaStreamSynchronize(mp->copy_stream);
  if (mp->aux_stream) cudaStreamSynchronize(mp->aux_stream);
  free_scratch(mp);
  if (mp->ev_geom) cudaEventDestroy(mp->ev_geom);
  if (mp->ev_rays) cudaEventDestroy(mp->ev_rays);
  if (mp->aux_stream) cudaStreamDestroy(mp->aux_stream);
  destroy_scan_graph(mp);
  destroy_batch_graph(mp);
  cudaFree(mp->d_pm2);
  cudaFree(mp->d_keys2);
  cudaFree(mp->tb2.records);
  cudaFree(mp->tb2.bucket_count);
  cudaFree(mp->tb2.bucket_offset);
  cudaFree(mp->tb2.bucket_cursor);
  cudaFree(mp->tb2.bucket_list);
  cudaFree(mp->tb2.heavy_list);
  cudaFree(mp->d_ring);
  cudaFree(mp->d_move);
  cudaFree(mp->d_batch_counters);
  cudaFreeHost(mp->h_batch_counters);
  for (cudaEvent_t e : mp->ev_copied)
    if (e) cudaEventDestroy(e);
  if (mp->copy_stream) cudaStreamDestroy(mp->copy_stream);
  cudaFree(mp->d_counters);
  cudaFree(mp->tb.bucket_count);
  cudaFree(mp->tb.bucket_offset);
  cudaFree(mp->tb.bucket_cursor);
  cudaFree(mp->tb.bucket_list);
  cudaFree(mp->tb.heavy_list);
  drain_stage_events(mp);
  for (cudaEvent_t e : mp->ev_pool) cudaEventDestroy(e);
  delete mp;
  return FDEM_OK;
}

fdem_status fdem_mapper_set_config(fdem_mapper* mp, const fdem_config* cfg) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && cfg, "null argument");
  FDEM_TRY(validate_config(cfg));
  DeviceGuard dg(mp->map->device);
  mp->cfg = *cfg;
  // setEstimatorType/setMappingMode re-create ElevationMapping (fastdem.cpp:28-38), whose
  // ctor adds whatever estimator layers are missing; existing data persists
  return ensure_estimator_layers(mp->map, mp->cfg);
}

fdem_status fdem_mapper_get_config(fdem_mapper* mp, fdem_config* out) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && out, "null argument");
  *out = mp->cfg;
  return FDEM_OK;
}

static fdem_status integrate_common(fdem_mapper* mp, const float* xyzw, const float* cov9,
                                    const float* intensity, const uint8_t* rgb, size_t n,
                                    const double* Tbs, const double* Twb) {
  FDEM_MAPPER_ALIVE(mp);
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
  FDEM_MAPPER_ALIVE(mp);
  FDEM_TRY(integrate_common(mp, xyzw, nullptr, intensity, rgb, n, Tbs, Twb));
  DeviceGuard dg(mp->map->device);
  return finish_scan(mp, stats);
}

fdem_status fdem_mapper_integrate_with_cov(fdem_mapper* mp, const float* xyzw, const float* cov9,
                                           const float* intensity, const uint8_t* rgb, size_t n,
                                           const double* Tbs, const double* Twb,
                                           fdem_scan_stats* stats) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(cov9 || n == 0, "cov9 is null");
  FDEM_TRY(integrate_common(mp, xyzw, cov9, intensity, rgb, n, Tbs, Twb));
  DeviceGuard dg(mp->map->device);
  return finish_scan(mp, stats);
}

static fdem_status enqueue_pointcloud2(fdem_mapper* mp, const uint8_t* data, size_t n,
                                       const fdem_pointcloud2_layout* lo, const double* Tbs,
                                       const double* Twb);

fdem_status fdem_mapper_integrate_pointcloud2(fdem_mapper* mp, const uint8_t* data, size_t n,
                                              const fdem_pointcloud2_layout* lo, const double* Tbs,
                                              const double* Twb, fdem_scan_stats* stats) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && lo && Tbs && Twb, "null argument");
  DeviceGuard dg(mp->map->device);
  // from_impl: empty message or no xyz fields -> empty cloud -> integrate() returns false
  if (n == 0 || lo->off_x < 0 || lo->off_y < 0 || lo->off_z < 0) {
    mp->last = fdem_scan_stats{};
    mp->last_had_work = false;
    if (stats) *stats = mp->last;
    return FDEM_OK;
  }
  FDEM_TRY(enqueue_pointcloud2(mp, data, n, lo, Tbs, Twb));
  return finish_scan(mp, stats);
}

fdem_status fdem_mapper_submit_pointcloud2(fdem_mapper* mp, const uint8_t* data, size_t n,
                                           const fdem_pointcloud2_layout* lo, const double* Tbs,
                                           const double* Twb, uint64_t* ticket) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && lo && Tbs && Twb && ticket, "null argument");
  FDEM_REQUIRE(n > 0 && lo->off_x >= 0 && lo->off_y >= 0 && lo->off_z >= 0,
               "submit needs a non-empty cloud with x/y/z fields (integrate() returns false otherwise)");
  DeviceGuard dg(mp->map->device);
  const uint64_t before = mp->map->seq;
  // never let the ring wrap over a result nobody has been able to collect yet
  if (before >= kResultRing)
    FDEM_CUDA_TRY(cudaEventSynchronize(mp->map->ev_scan[(before - kResultRing + 1) % kResultRing]));
  FDEM_TRY(enqueue_pointcloud2(mp, data, n, lo, Tbs, Twb));
  *ticket = before;
  return FDEM_OK;
}

static fdem_status enqueue_pointcloud2(fdem_mapper* mp, const uint8_t* data, size_t n,
                                       const fdem_pointcloud2_layout* lo, const double* Tbs,
                                       const double* Twb) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(data, "data is null");
  FDEM_REQUIRE(lo->point_step >= 12 && (lo->point_step & 3) == 0, "point_step must be a multiple of 4");
  auto fits = [&](int32_t off, int32_t size) { return off >= 0 && off + size <= static_cast<int32_t>(lo->point_step); };
  FDEM_REQUIRE(fits(lo->off_x, 4) && fits(lo->off_y, 4) && fits(lo->off_z, 4) &&
                   !((lo->off_x | lo->off_y | lo->off_z) & 3), "x/y/z offsets out of range or unaligned");
  if (lo->off_intensity >= 0) {
    const int32_t t = lo->intensity_type;
    const int32_t sz = t == 2 ? 1 : t == 4 ? 2 : t == 7 ? 4 : t == 8 ? 8 : 0;
    // unknown datatypes read as 0.0f in the reference (readIntensity default branch)
    if (sz) FDEM_REQUIRE(fits(lo->off_intensity, sz) && (lo->off_intensity % (sz >= 4 ? 4 : sz)) == 0,
                         "intensity offset out of range or unaligned");
  }
  if (lo->off_rgb >= 0)
    FDEM_REQUIRE(fits(lo->off_rgb, 4) && (lo->off_rgb & 3) == 0, "rgb offset out of range or unaligned");
  ScanInputs in{};
  in.raw = data;
  in.layout = lo;
  in.n = n;
  in.input_frame = INPUT_SENSOR_FRAME;
  in.Tbs = Tbs;
  in.Twb = Twb;
  in.robot_x = Twb[12];
  in.robot_y = Twb[13];
  return enqueue_scan(mp, in);
}

fdem_status fdem_mapper_integrate_async(fdem_mapper* mp, const float* xyzw,
                                        const float* intensity, const uint8_t* rgb, size_t n,
                                        const double* Tbs, const double* Twb) {
  FDEM_MAPPER_ALIVE(mp);
  return integrate_common(mp, xyzw, nullptr, intensity, rgb, n, Tbs, Twb);
}

fdem_status fdem_mapper_submit(fdem_mapper* mp, const float* xyzw, const float* intensity,
                               const uint8_t* rgb, size_t n, const double* Tbs, const double* Twb,
                               uint64_t* ticket) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && ticket, "null argument");
  FDEM_REQUIRE(n > 0, "submit needs a non-empty cloud (integrate() returns false on empty input)");
  const uint64_t before = mp->map->seq;
  // never let the ring wrap over a result nobody has been able to collect yet
  if (before >= kResultRing) {
    DeviceGuard dg(mp->map->device);
    FDEM_CUDA_TRY(cudaEventSynchronize(mp->map->ev_scan[(before - kResultRing + 1) % kResultRing]));
  }
  FDEM_TRY(integrate_common(mp, xyzw, nullptr, intensity, rgb, n, Tbs, Twb));
  *ticket = before;
  return FDEM_OK;
}

fdem_status fdem_mapper_collect(fdem_mapper* mp, uint64_t ticket, fdem_scan_stats* stats) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && stats, "null argument");
  fdem_map* m = mp->map;
  FDEM_REQUIRE(ticket < m->seq, "unknown ticket");
  FDEM_REQUIRE(ticket + kResultRing > m->seq, "ticket expired: its result slot has been reused");
  DeviceGuard dg(m->device);
  FDEM_CUDA_TRY(cudaEventSynchronize(m->ev_scan[ticket % kResultRing]));
  const ScanResult& r = m->h_result[ticket % kResultRing];
  stats->n_input = 0;  // not retained per ticket
  stats->n_kept = r.counters[CNT_KEPT];
  stats->n_cells = r.counters[CNT_CELLS];
  stats->n_voxels = r.counters[CNT_VOXELS];
  stats->integrated = r.counters[CNT_KEPT] > 0 ? 1 : 0;
  stats->voxel_box_violations = static_cast<int32_t>(r.counters[CNT_VOX_VIOLATION]);
  if (ticket + 1 == m->seq) {  // newest scan: its committed geometry is the map's geometry
    m->geom = r.state.geom;
    m->geom_stale = false;
    apply_state_flags(m, r.state.flags);
  }
  adapt_bucket_shape(mp, r.counters[CNT_BUCKETS], r.counters[CNT_CELLS]);
  return FDEM_OK;
}

fdem_status fdem_mapper_integrate_batch(fdem_mapper* mp, int32_t n_scans, const float* const* xyzw,
                                        const float* const* intensity, const uint8_t* const* rgb,
                                        const size_t* n_points, const double* Tbs, const double* Twb,
                                        fdem_scan_stats* stats) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && xyzw && n_points && Tbs && Twb, "null argument");
  FDEM_REQUIRE(n_scans >= 1 && n_scans <= kMaxBatch, "a batch holds 1 to 16 scans");
  fdem_map* m = mp->map;
  DeviceGuard dg(m->device);
  const bool overlapped = mp->use_tile && mp->use_graph && !mp->stage_timing && !mp->cfg.raycasting_enabled;
  bool on_device = true;
  for (int i = 0; i < n_scans; ++i) {
    FDEM_REQUIRE(n_points[i] > 0 && xyzw[i], "every scan of a batch must be non-empty");
    on_device = on_device && is_device_pointer(xyzw[i]) && (!intensity || !intensity[i] || is_device_pointer(intensity[i])) &&
                (!rgb || !rgb[i] || is_device_pointer(rgb[i]));
  }
  if (!overlapped || !on_device) {
    // same results, one scan after the other (raycasting, the global-sort path and host input
    // buffers keep the per-scan pipeline)
    for (int i = 0; i < n_scans; ++i) {
      FDEM_TRY(integrate_common(mp, xyzw[i], nullptr, intensity ? intensity[i] : nullptr, rgb ? rgb[i] : nullptr,
                                n_points[i], Tbs + 16 * i, Twb + 16 * i));
      FDEM_TRY(finish_scan(mp, stats ? stats + i : nullptr));
    }
    return FDEM_OK;
  }
  cudaStream_t s = m->stream;
  size_t n_max = 0;
  for (int i = 0; i < n_scans; ++i) n_max = std::max(n_max, n_points[i]);
  FDEM_TRY(ensure_capacity(mp, n_max));
  FDEM_TRY(ensure_batch_scratch(mp));
  if (mp->tile_dirty) {
    // after a failed scan the bucket scratch may hold leftovers: the primary set is re-zeroed by
    // the scan builder below, the second set here
    FDEM_CUDA_TRY(cudaMemsetAsync(mp->tb2.bucket_count, 0, mp->max_buckets * sizeof(uint32_t), s));
    FDEM_CUDA_TRY(cudaMemsetAsync(mp->tb2.bucket_cursor, 0, mp->max_buckets * sizeof(uint32_t), s));
  }
  if (!mp->bg) mp->bg = new (std::nothrow) BatchGraphState[2]();
  mp->bg_flip ^= 1;
  if (!mp->bg) return set_error(FDEM_ERR_OUT_OF_MEMORY, "host allocation failed");
  // Result slots are reused every kResultRing scans.  A batch hands out no tickets, so nobody
  // can still want an older slot's content: no wait here (the host keeps queueing batches while
  // the device works; fdem_mapper_wait reads the newest slot after a stream sync).
  const uint64_t ticket0 = m->seq;
  if (m->obstacle_full_clear) {
    or_state_flags_kernel<<<1, 1, 0, s>>>(m->d_state, SF_OBSTACLE_DIRTY);
    ++m->lc.mine;
    m->obstacle_full_clear = false;
  }
  for (int i = 0; i < n_scans; ++i) {
    ScanInputs in{};
    in.xyzw = xyzw[i];
    in.intensity = intensity ? intensity[i] : nullptr;
    in.rgb = rgb ? rgb[i] : nullptr;
    in.n = n_points[i];
    in.input_frame = INPUT_SENSOR_FRAME;
    in.Tbs = Tbs + 16 * i;
    in.Twb = Twb + 16 * i;
    in.robot_x = in.Twb[12];
    in.robot_y = in.Twb[13];
    in.known_device = true;
    BatchScan& b = mp->bg[mp->bg_flip].scan[i];
    FDEM_TRY(enqueue_scan(mp, in, &b.L, ticket0 + i));
    b.n = static_cast<uint32_t>(n_points[i]);
    retarget_for_batch(mp, i, b);
    b.L.pub.enabled = i == n_scans - 1 ? 1 : 0;
  }
  fdem_status st = launch_batch_graph(mp, n_scans, s);
  if (st != FDEM_OK) {
    mp->tile_dirty = true;
    mp->counters_dirty = true;
    return st;
  }
  for (int i = 0; i < n_scans; ++i)
    FDEM_CUDA_TRY(cudaEventRecord(m->ev_scan[(ticket0 + i) % kResultRing], s));
  m->lc.mine += static_cast<int64_t>(mp->tb.job_counter == CNT_HEAVY ? BN_COUNT : BN_COUNT - 1) * n_scans;
  m->seq = ticket0 + n_scans;
  mp->last_ticket = ticket0 + n_scans - 1;
  m->geom_stale = true;
  mp->last_n = static_cast<uint32_t>(n_points[n_scans - 1]);
  mp->last_raw = false;
  mp->last_had_work = true;
  mp->pending = true;
  mp->last_batch_n = n_scans;
  mp->last_batch_ticket0 = ticket0;
  for (int i = 0; i < n_scans; ++i) mp->last_batch_points[i] = n_points[i];
  if (!stats) return FDEM_OK;   // queued: fdem_mapper_wait(), then fdem_mapper_last_batch_stats()
  FDEM_CUDA_TRY(cudaStreamSynchronize(s));
  FDEM_TRY(fdem_mapper_last_batch_stats(mp, stats, n_scans));
  return finish_scan(mp, nullptr);
}

fdem_status fdem_mapper_last_batch_stats(fdem_mapper* mp, fdem_scan_stats* stats, int32_t n_scans) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && stats, "null argument");
  FDEM_REQUIRE(n_scans == mp->last_batch_n && n_scans > 0, "n_scans must equal the size of the last batch");
  fdem_map* m = mp->map;
  DeviceGuard dg(m->device);
  FDEM_REQUIRE(mp->last_batch_ticket0 + n_scans == m->seq, "other scans were integrated after the batch");
  FDEM_CUDA_TRY(cudaStreamSynchronize(m->stream));
  const ScanResult& last = m->h_result[(mp->last_batch_ticket0 + n_scans - 1) % kResultRing];
  for (int i = 0; i < n_scans; ++i) {
    const uint32_t* c = i == n_scans - 1 ? last.counters : mp->h_batch_counters + static_cast<size_t>(i) * CNT_COUNT;
    stats[i].n_input = static_cast<int64_t>(mp->last_batch_points[i]);
    stats[i].n_kept = c[CNT_KEPT];
    stats[i].n_cells = c[CNT_CELLS];
    stats[i].n_voxels = 0;
    stats[i].integrated = c[CNT_KEPT] > 0 ? 1 : 0;
    stats[i].voxel_box_violations = 0;
  }
  return FDEM_OK;
}

fdem_status fdem_mapper_wait(fdem_mapper* mp, fdem_scan_stats* stats) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp, "null mapper");
  DeviceGuard dg(mp->map->device);
  return finish_scan(mp, stats);
}

fdem_status fdem_mapper_update(fdem_mapper* mp, const float* xyzw, const float* var_z,
                               const float* intensity, const uint8_t* rgb, size_t n,
                               double robot_x, double robot_y, fdem_scan_stats* stats) {
  FDEM_MAPPER_ALIVE(mp);
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
  FDEM_MAPPER_ALIVE(mp);
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
  FDEM_MAPPER_ALIVE(mp);
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

#ifdef FDEM_PROBES
// tuning probes (tools/phase_probe.py): exported only by a -DFDEM_PROBES build, not declared in
// the public header
__attribute__((visibility("default"))) fdem_status fdem_mapper_debug_cta_times(fdem_mapper* mp, uint64_t* out1024) {
  FDEM_REQUIRE(mp && out1024, "null argument");
  DeviceGuard dg(mp->map->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(mp->map->stream));
  FDEM_CUDA_TRY(static_cast<cudaError_t>(
      tile_estimate_debug_cta_ns(reinterpret_cast<unsigned long long*>(out1024))));
  return FDEM_OK;
}

__attribute__((visibility("default"))) fdem_status fdem_mapper_debug_phase_clocks(fdem_mapper* mp, int64_t* out16) {
  FDEM_REQUIRE(mp && out16, "null argument");
  DeviceGuard dg(mp->map->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(mp->map->stream));
  long long tmp[16];
  FDEM_CUDA_TRY(static_cast<cudaError_t>(tile_estimate_debug_clocks(tmp)));
  for (int i = 0; i < 16; ++i) out16[i] = tmp[i];
  return FDEM_OK;
}

#endif

fdem_status fdem_mapper_set_cell_sort(fdem_mapper* mp, int32_t mode) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp, "null mapper");
  FDEM_REQUIRE(mode == FDEM_CELL_SORT_TILE || mode == FDEM_CELL_SORT_GLOBAL, "bad cell sort mode");
  DeviceGuard dg(mp->map->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(mp->map->stream));
  mp->use_tile = mode == FDEM_CELL_SORT_TILE;
  mp->tile_dirty = true;
  return FDEM_OK;
}

fdem_status fdem_mapper_set_voxel_sort(fdem_mapper* mp, int32_t mode) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp, "null mapper");
  FDEM_REQUIRE(mode == FDEM_VOXEL_SORT_MSD || mode == FDEM_VOXEL_SORT_LIBRARY, "bad voxel sort mode");
  DeviceGuard dg(mp->map->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(mp->map->stream));
  if (mp->aux_stream) FDEM_CUDA_TRY(cudaStreamSynchronize(mp->aux_stream));
  mp->voxel_sort_library = mode == FDEM_VOXEL_SORT_LIBRARY;
  return FDEM_OK;
}

fdem_status fdem_mapper_set_stage_timing(fdem_mapper* mp, int32_t enabled) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp, "null mapper");
  DeviceGuard dg(mp->map->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(mp->map->stream));
  drain_stage_events(mp);
  mp->stage_timing = enabled != 0;
  for (double& v : mp->stage_ms) v = 0.0;
  mp->stage_scans = 0;
  return FDEM_OK;
}

fdem_status fdem_mapper_stage_times(fdem_mapper* mp, double* ms, int64_t* scans) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && ms && scans, "null argument");
  DeviceGuard dg(mp->map->device);
  FDEM_CUDA_TRY(cudaStreamSynchronize(mp->map->stream));
  drain_stage_events(mp);
  for (int i = 0; i < FDEM_STAGE_COUNT; ++i) {
    ms[i] = mp->stage_ms[i];
    mp->stage_ms[i] = 0.0;
  }
  *scans = mp->stage_scans;
  mp->stage_scans = 0;
  return FDEM_OK;
}

fdem_status fdem_mapper_library_launch_count(fdem_mapper* mp, int64_t* launches) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && launches, "null argument");
  *launches = mp->map->lc.library;
  return FDEM_OK;
}

fdem_status fdem_mapper_launch_count(fdem_mapper* mp, int64_t* launches) {
  FDEM_MAPPER_ALIVE(mp);
  FDEM_REQUIRE(mp && launches, "null argument");
  *launches = mp->map->lc.mine;
  return FDEM_OK;
}

}  // extern "C"

// raycasting / voxel / inpainting entry points
#include "capi_post.inc"
// multi-GPU GLOBAL map
#include "capi_shard.inc"
