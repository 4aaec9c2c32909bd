// device_types.h — parameter blocks passed by value to the kernels, and the launcher
// prototypes capi.cu calls.  POD only; no torch / Eigen types.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "grid_geom.h"

namespace fdem {

// per-scan counters in device memory (zeroed at the start of every scan)
enum Counter : int {
  CNT_KEPT = 0,      // points surviving cropRange + cropZ
  CNT_INSIDE = 1,    // kept points that fall inside the map (= valid sort keys)
  CNT_CELLS = 2,     // touched cells (segments)
  CNT_VOXELS = 3,    // voxelGrid(ANY) representatives
  CNT_FINITE = 4,    // PointCloud2 ingest: points with finite x, y, z (= cloud.size() after from())
  CNT_RC_SKIP = 5,   // 1 when raycasting preconditions failed (sensor outside map)
  CNT_VOX_VIOLATION = 6,  // points outside the host-predicted voxel box (must stay 0)
  CNT_RAYS = 7,        // raycasting: rays in the dense list (downward representatives)
  CNT_REC_SLOTS = 8,   // tile path: record slots handed out to bucket segments
  CNT_BUCKETS = 9,     // tile path: non-empty buckets this scan
  CNT_WORK = 10,       // tile path: K3t's dynamic work counter
  CNT_RAY_WORK = 11,   // raycasting: next unfetched ray bundle
  CNT_HEAVY = 12,      // tile path: buckets the warp-per-bucket kernel left to the CTA-per-bucket kernel
  CNT_COUNT = 16
};

// ── tile path: 2-level sort-by-cell ───────────────────────────────────────────
// L1 = radix partition of the scan into buckets of 2^kBucketBits consecutive cell keys
// (global, one pass: histogram in K1, segment allocation in K2, scatter kernel);
// L2 = per-bucket sort + warp-segmented reduce in shared memory (K3t).
// The bucket size is a per-mapper choice (bucket_bits = 9 or 10): K3t is compiled for both.

// one run of same-cell points, pre-reduced by the scatter kernel (fields = CellObs)
struct alignas(32) CellRecord {
  uint32_t lkey;  // cell index inside its bucket
  float mz, mv;
  uint32_t mi;
  float xz, it;
  uint32_t fi, li;
};
static_assert(sizeof(CellRecord) == 32, "CellRecord must be 32 bytes (TMA bulk-copy unit)");

struct TileBuffers {
  uint32_t* bucket_count;   // [n_buckets] points per bucket (K1); K3t re-zeroes what it consumed
  uint32_t* bucket_offset;  // [n_buckets] first record slot of the bucket's segment (K2)
  uint32_t* bucket_cursor;  // [n_buckets] records written so far (scatter); K3t re-zeroes
  uint4* bucket_list;       // [n_buckets] non-empty buckets: {bucket, first slot, points, 0} (K2)
  CellRecord* records;      // [capacity in points]
  uint32_t n_buckets;
  uint32_t bucket_bits;     // cells per bucket = 1 << bucket_bits
  // K3t's work list: bucket_list with counters[CNT_BUCKETS] entries (K2), or — after the light
  // pass has taken the small buckets — heavy_list with counters[CNT_HEAVY] entries
  uint4* heavy_list;        // [n_buckets]
  uint32_t job_counter;     // CNT_BUCKETS (0 means that too) or CNT_HEAVY
};

// State that survives from scan to scan and is decided on the device (so a stream of
// scans needs no host round trip).  Double buffered: scan s reads [s&1], the commit
// kernel writes [(s+1)&1].
struct DeviceState {
  GridGeom geom;
  uint32_t touched_count;  // valid entries in the touched-key list of the last scan with >=1 cell
  // sticky facts only the device learns, carried from scan to scan (StateFlag bits): which
  // lazily created layers the reference would have created by now, and whether the obstacle
  // layer was written behind the mapper's back (needs the reference's whole-layer clear)
  uint32_t flags;
};
enum StateFlag : uint32_t {
  SF_INTENSITY = 1u,      // a scan carrying intensity produced >= 1 cell (elevation_mapping.cpp:154-156)
  SF_COLOR = 2u,          // a scan carrying colour produced >= 1 cell (:168-170)
  SF_RAYCAST = 4u,        // applyRaycasting got past its guards once (raycasting.cpp:221-240)
  SF_OBSTACLE_DIRTY = 8u  // obstacle layer edited by the caller: next observing scan clears all of it
};

enum InputFrame : int { INPUT_SENSOR_FRAME = 0, INPUT_MAP_FRAME = 1 };

// ── multi-GPU GLOBAL map: row stripes ─────────────────────────────────
// One logical map, `world` stripes of consecutive logical rows, one per rank / GPU: the first
// `extra` stripes hold base + 1 rows, the others base (fastdem_b200/sharded.py stripe_bounds).
// A scan is integrated in two halves.  FRONT (every rank, on its 1/world slice of the scan's
// points, read in place from the ingest GPU over NVLink): K1 + bucket partition, with
// owner-major cell keys  key = stripe * stride + (col * rows_of_stripe + row_in_stripe), so a
// 1024-key bucket never straddles two stripes; the records stay in the source rank's memory.
// BACK (every rank, for the buckets of its own stripe): pull the bucket's record pieces from
// all `world` sources over NVLink (TMA bulk reads of peer memory) and run K3t.  Two flags per
// pair of ranks order the halves (ready: source -> owner, consumed: owner -> source); there is
// no collective and no host round trip on the data path.
constexpr int kMaxShards = 8;

struct ShardGeom {
  int32_t world, rank;     // world <= 1: unsharded
  int32_t base, extra;     // rows / world, rows % world
  uint32_t stride;         // key stride between stripes (cells of the largest stripe, rounded up to 1024)
};

// stripe that owns logical row `row`, its first row and its row count
FDEM_HD int32_t shard_of_row(const ShardGeom& sg, int32_t row, int32_t& row_begin, int32_t& rows_local) {
  const int32_t big = sg.extra * (sg.base + 1);
  if (row < big) {
    const int32_t d = row / (sg.base + 1);
    row_begin = d * (sg.base + 1);
    rows_local = sg.base + 1;
    return d;
  }
  const int32_t d = sg.extra + (row - big) / sg.base;
  row_begin = big + (d - sg.extra) * sg.base;
  rows_local = sg.base;
  return d;
}

// what a source rank publishes for the owners, and the flags of the two-way handshake; lives at
// the start of every rank's exchange arena (peer-mapped by all other ranks)
// scans in flight between a rank's front half and the owners' back halves: every per-scan buffer of
// the exchange (tables, record areas, counters, flags' payload) exists kShardDepth times, indexed
// by seq % kShardDepth; a source may run its front half up to kShardDepth scans ahead of the
// slowest owner.  Measured on 2 x B200 (C5): depth 4 = 14.81 K scans/s, depth 2 = 14.76 K — the
// job is bound by the busiest owner's serial back-half chain, not by the buffering depth; 2 is
// the depth validated at N = 2, 4 and 8.
constexpr uint32_t kShardDepth = 2;

struct ShardHeader {
  uint32_t ready[kMaxShards];        // [s] written by source s: front half of scan #ready[s] is complete
  uint32_t consumed[kMaxShards];     // [d] written by owner d: it has finished reading scan #consumed[d]
  uint32_t inside[kShardDepth][kMaxShards];  // [seq % depth][s]: points of source s's slice inside the map
  uint32_t load[kShardDepth][kMaxShards];    // [seq % depth][d] written by owner d with consumed[d]: cells of
                                             // its stripe that scan touched (the slice split of scan
                                             // seq + depth is derived from it)
  uint32_t _pad[16];
};

// A rank's slice of the scan for the front half, decided ON THE DEVICE by shard_begin_kernel from
// the owners' loads of kShardDepth scans ago (identical on every rank): ranks whose stripe got most of
// the cells — and so most of the back half — bin fewer points.
struct ShardSlice {
  uint32_t begin, count;
};

#if defined(__CUDACC__)
#define FDEM_HOST_DEVICE __host__ __device__
#else
#define FDEM_HOST_DEVICE
#endif
// The slice of an n_scan-point scan that rank `rank` bins, given the cells each stripe's owner
// touched (`load`).  Integer arithmetic only, identical on every rank and on the host
// (fdem_shard_slice_plan).  A rank's work is (its share of the points) + w * (its share of the
// cells), both as fractions in 1/65536, w = back_weight / 256 = the cost of a whole back half in
// units of a whole front half; the shares of the points level that sum (water-filling): a rank
// that owns most of the cells bins few points or none, the ranks with idle stripes bin the rest.
// All loads zero: equal slices.  Boundaries are warp-aligned; slices tile [0, n_scan) in rank order.
FDEM_HOST_DEVICE inline ShardSlice shard_slice_plan(const uint32_t* load, int world, int rank, uint32_t n_scan,
                                                    uint32_t back_weight) {
  uint32_t q[kMaxShards], share[kMaxShards];
  uint64_t total = 0;
  for (int d = 0; d < world; ++d) total += load[d];
  constexpr uint32_t ONE = 65536u;
  if (total == 0) {
    for (int d = 0; d < world; ++d) share[d] = ONE / static_cast<uint32_t>(world);
  } else {
    for (int d = 0; d < world; ++d)
      q[d] = static_cast<uint32_t>((static_cast<uint64_t>(load[d]) * ONE / total * back_weight) >> 8);
    uint32_t active = (1u << world) - 1u, level = 0;
    for (int it = 0; it < world; ++it) {
      uint32_t sum = ONE, cnt = 0;
      for (int d = 0; d < world; ++d)
        if (active >> d & 1u) { sum += q[d]; ++cnt; }
      level = sum / cnt;
      uint32_t drop = 0;
      for (int d = 0; d < world; ++d)
        if ((active >> d & 1u) && q[d] > level) drop |= 1u << d;
      if (!drop) break;
      active &= ~drop;   // (the rank with the smallest load always stays)
    }
    for (int d = 0; d < world; ++d) share[d] = (active >> d & 1u) ? level - q[d] : 0u;
  }
  // cumulative shares scaled to the scan; the last rank with a share takes the rounding remainder
  uint32_t acc = 0, total_share = 0;
  for (int d = 0; d < world; ++d) total_share += share[d];
  uint32_t b0 = 0, b1 = 0;
  for (int d = 0; d <= rank; ++d) {
    b0 = b1;
    acc += share[d];
    b1 = acc == total_share ? n_scan
                            : static_cast<uint32_t>(static_cast<uint64_t>(n_scan) * acc / total_share) & ~31u;
  }
  return ShardSlice{b0, b1 - b0};
}

// one non-empty bucket of this rank's stripe: where its records lie in every source's arena
struct ShardJob {
  uint32_t bucket;                   // bucket index inside the stripe
  uint32_t total;                    // records over all sources
  uint32_t off[kMaxShards];          // first record slot in source s's record buffer
  uint32_t cnt[kMaxShards];
};

struct ShardFrontArgs {              // source side, after the scatter: publish to the owners
  ShardHeader* peer_hdr[kMaxShards];
  uint32_t seq;
  int32_t world, rank;
};

struct ShardBackArgs {               // owner side
  ShardHeader* hdr;                            // this rank's header (flags written by the peers)
  ShardHeader* peer_hdr[kMaxShards];           // consumed[] is written back to every source
  const uint32_t* peer_cursor[kMaxShards];     // source s's record count per (global) bucket, this scan's parity
  const uint32_t* peer_offset[kMaxShards];     // source s's first record slot per bucket
  const CellRecord* peer_records[kMaxShards];  // source s's record buffer, this scan's parity
  ShardJob* jobs;                              // [buckets per stripe]
  ShardJob* heavy_jobs;                        // [buckets per stripe] left over by the light pass
  uint32_t job_counter;                        // which counter holds the length of the list K3t walks
  uint32_t seq;
  int32_t world, rank;
  uint32_t bps;                                // buckets per stripe (= stride >> 10)
  // map-side bookkeeping the back half owns (obstacle reset, touched-list hand-over, state flags)
  const DeviceState* st_cur;
  DeviceState* st_out;
  float* obstacle;
  const uint32_t* touched_keys;
  uint32_t invalid_key;
  uint32_t flags_if_cells;
  size_t obstacle_cells;
  const uint32_t* front_counters;  // this rank's front-half counters of the scan (kept / inside of its slice)
};

// everything K1 needs for one scan (fastdem::Config + the two transforms, pre-cast on the
// host exactly as the reference casts them: Isometry3d::matrix().cast<float>(),
// nanopcl/core/transform.hpp:68-82; (T_world_base*T_base_sensor).rotation().cast<float>(),
// fastdem/src/fastdem.cpp:182-183)
struct PreprocessParams {
  const float4* xyzw;
  const float* intensity;  // unused by K1; carried for symmetry
  const float* cov9;       // optional caller-provided sensor-frame covariances (N x 9, col-major)
  const float* var_z;      // optional (map-frame input): cloud.covariance(i)(2,2)
  uint32_t n;
  const ShardSlice* slice;  // multi-GPU front half: bin points [begin, begin + count) of the scan
                            // (read on the device; n bounds the grid)
  int32_t input_frame;
  float T1[16];  // T_base_sensor as Matrix4f, column-major
  float T2[16];  // T_world_base as Matrix4f, column-major
  float R[9];    // rotation of T_world_base*T_base_sensor, column-major
  double robot_x, robot_y;
  float z_min, z_max, range_min_sq, range_max_sq;
  int32_t sensor_type;
  float lidar_range_noise, lidar_angular_noise;  // already fabs()'d
  float rgbd_a, rgbd_b, rgbd_c, rgbd_k;
  float constant_variance;  // uncertainty^2
  int32_t local_mode;
  uint32_t invalid_key;  // = number of cells in this handle's slab
  uint32_t* bucket_count;  // tile path: per-bucket point histogram (null on the global-sort path)
  uint32_t bucket_bits;
  int32_t write_vals;      // global-sort path needs vals[i] = i
  // sensor_msgs/PointCloud2 ingest (nanopcl::from(msg), nanopcl/bridge/ros/impl.hpp:180-270):
  // when `raw` is set, point i is read from raw + i*point_step at the field offsets, points
  // with a non-finite coordinate are dropped, intensity / packed rgb are unpacked into the
  // SoA scratch the later kernels read
  const uint8_t* raw;
  uint32_t point_step;
  int32_t off_x, off_y, off_z, off_intensity, off_rgb;
  int32_t intensity_type;  // PointField datatype: 2 UINT8, 4 UINT16, 7 FLOAT32, 8 FLOAT64
  float* out_intensity;
  uint8_t* out_rgb;
  int32_t raw_vec16;  // 16-byte points with x, y, z at 0, 4, 8 and one 4-byte channel at 12, base 16-B aligned
  ShardGeom shard;    // world > 1: owner-major keys for every stripe (multi-GPU front half)
};

// pointers to every layer the estimator kernel reads or writes (null when absent)
struct EstLayers {
  float *elevation, *elevation_min, *elevation_max;
  float *variance, *n_points, *upper_bound, *lower_bound;
  float *obstacle, *intensity, *color;
  float *kalman_p, *sample_mean, *sample_m2;
  float* p2_q[5];
  float* p2_n[5];
};

struct EstimateParams {
  const uint32_t* sorted_keys;
  const uint32_t* sorted_vals;
  const float4* pm;        // K1 output: map-frame x, y, z, var_z per input point
  const float* intensity;  // input channel (null when the cloud has none)
  const uint8_t* rgb;      // input channel (null when the cloud has none)
  uint32_t* touched_keys;  // out: key at segment heads, invalid_key elsewhere
  float* touched_minz;     // out (optional): per-scan min_z at segment heads
  uint32_t n_sorted;
  uint32_t invalid_key;
  int32_t estimation_type;
  float kalman_min_variance, kalman_max_variance, kalman_process_noise;
  float p2_dn[5];  // already clamped + monotonised (quantile_estimation.hpp:83-94)
  int32_t p2_marker;
  float p2_max_sample_count;
  EstLayers L;
};

constexpr int kMaxLayers = 48;
struct LayerTable {
  float* ptr[kMaxLayers];
  int32_t count;
  int32_t basic[3];  // indices of elevation, elevation_min, elevation_max
};

// what a deferred commit hands to the back prologue of the same scan (batched integration)
struct MoveRecord {
  MoveResult mr;
  GridGeom g_new;
};

struct CommitParams {
  double robot_x, robot_y;
  int32_t local_mode;
  int32_t clear_policy;
  uint32_t invalid_key;
  float* obstacle;
  const uint32_t* touched_keys;
  // batched integration: the commit only computes the new geometry (so the NEXT scan's K1 can
  // start) and records the move; every write to map layers waits for back_prologue_kernel,
  // which runs after the previous scan's estimator
  int32_t defer;
  MoveRecord* move_out;
  int32_t tile_path;  // 1: also hand out bucket segments (TileBuffers) and let K3t count cells
  TileBuffers tb;
  // StateFlag bookkeeping: bits to set when this scan produces >= 1 cell (its channels), and the
  // raycasting guards (enabled; sensor origin, tested against the committed geometry)
  uint32_t flags_if_cells;
  int32_t raycast;
  double rc_origin_x, rc_origin_y;
  size_t obstacle_cells;  // cells in the obstacle layer (whole-layer clear when SF_OBSTACLE_DIRTY)
};

struct ScatterParams {
  const uint32_t* keys;
  const float4* pm;
  const float* intensity;
  uint32_t n;
  uint32_t invalid_key;
  TileBuffers tb;
  // the scatter grid also resets the last observing scan's obstacle cells (tile path)
  const uint32_t* counters;
  const DeviceState* st_cur;
  float* obstacle;
  const uint32_t* touched_keys;
  size_t obstacle_cells;
  uint32_t index_base;  // point index of element 0
  const ShardSlice* slice;  // multi-GPU: the slice K1 binned (overrides n / index_base, offsets intensity)
  // multi-GPU: records are PUSHED — a record of a bucket of stripe d goes straight into this
  // source's area of owner d's arena (peer-mapped; NVLink stores), so the owners' back halves
  // read local memory only.  owner_bps = buckets per stripe, 0 on a single GPU (tb.records).
  CellRecord* owner_records[kMaxShards];
  uint32_t owner_bps;
};

// back prologue of a scan in a batch: the map writes the commit / scatter kernels do in the
// one-scan pipeline (obstacle reset, LOCAL-mode move clearing, touched-count hand-over)
struct BackParams {
  const uint32_t* counters;
  const DeviceState* st_cur;   // state published by the previous scan (touched list length)
  DeviceState* st_out;         // this scan's state slot
  const MoveRecord* move;
  float* obstacle;
  const uint32_t* touched_keys;
  uint32_t invalid_key;
  int32_t clear_policy;
  size_t obstacle_cells;
};

// what K3t's last CTA needs to end the scan (publish); all null/0 when a separate
// publish_kernel launch follows instead (raycasting enabled)
struct PublishArgs {
  DeviceState* st_cur;
  uint32_t* host_out;
  int32_t enabled;
};

struct RaycastParams {
  float origin[3];
  float height_conflict_threshold, log_odds_observed, log_odds_ghost, log_odds_max,
      clear_threshold;
  float* elevation;
  float* raycasting;    // per-scan min ray height (NaN = not traversed)
  float* logodds;       // _visibility_logodds
  float* ghost_removal;
  uint32_t* ray_min_enc;  // scratch, one u32 per LOGICAL cell: bits of the lowest ray's (height - sensor z) <= -0; 0 = none
  uint32_t* hits;         // scratch, one u32 per cell: observed-evidence hit counts
  int32_t tune_ld_cg;     // far-field reads bypass L1 (always fresh) instead of L1-cached (maybe stale)
  int32_t tune_elect;     // same-cell lanes of a bundle elect one atomic (1: match.any, 2: uniform-cell fast path)
  int32_t seg_len;        // DDA steps per segment task
  int32_t tune_az_mask;   // azimuth bins actually used for the ordering key (mask on the 1024-bin index)
};

// ── launchers (kernels.cu) ───────────────────────────────────────────────────
struct LaunchCounter {
  int64_t mine = 0;     // kernels written in this repo
  int64_t library = 0;  // CUB kernels (radix sort passes)
};

void launch_fill(float* dst, size_t n, float v, cudaStream_t s, LaunchCounter& lc);
void launch_fill_u32(uint32_t* dst, size_t n, uint32_t v, cudaStream_t s, LaunchCounter& lc);
void launch_any_not_nan(const float* src, size_t n, uint32_t* flag, cudaStream_t s,
                        LaunchCounter& lc);
void launch_clear_cell(const LayerTable& lt, int64_t lin, cudaStream_t s, LaunchCounter& lc);
void launch_preprocess_bin(const PreprocessParams& p, const DeviceState* st_in, uint32_t* counters,
                           float4* pm, uint32_t* keys, uint32_t* vals, cudaStream_t s,
                           LaunchCounter& lc);
void launch_commit(const CommitParams& p, const DeviceState* st_in, DeviceState* st_out,
                   uint32_t* counters, const LayerTable& lt, cudaStream_t s, LaunchCounter& lc);
void launch_segreduce_estimate(const EstimateParams& p, const uint32_t* counters_ro,
                               uint32_t* counters, cudaStream_t s, LaunchCounter& lc);
// host_out: CNT_COUNT counter words followed by the words of DeviceState (mapped pinned memory)
void launch_publish(uint32_t* counters, DeviceState* st_cur, const DeviceState* st_next,
                    uint32_t* host_out, cudaStream_t s, LaunchCounter& lc);
// kernel entry + launch shape, for cudaGraphAddKernelNode / cudaLaunchKernel in capi.cu
struct KernelDesc {
  const void* func;
  dim3 grid, block;
  size_t smem;
};
KernelDesc desc_preprocess_bin(uint32_t n);
KernelDesc desc_commit();
KernelDesc desc_back_prologue();
KernelDesc desc_publish();
KernelDesc desc_scatter_records(uint32_t n);
KernelDesc desc_tile_estimate(uint32_t n_buckets, uint32_t bucket_bits);
KernelDesc desc_shard_begin();
KernelDesc desc_shard_alloc(int world);
KernelDesc desc_shard_publish_front();
KernelDesc desc_shard_gather();
KernelDesc desc_tile_estimate_shard(uint32_t bps);
KernelDesc desc_tile_estimate_light();
KernelDesc desc_tile_estimate_light_shard();
// tile path (kernels_tile.cu)
void launch_scatter_records(const ScatterParams& p, cudaStream_t s, LaunchCounter& lc);
void launch_tile_estimate(const EstimateParams& p, const TileBuffers& tb, uint32_t* counters,
                          DeviceState* st_out, const PublishArgs& pub, cudaStream_t s,
                          LaunchCounter& lc);
#ifdef FDEM_PROBES
int tile_estimate_debug_clocks(long long* out16);             // CTA 0 phase clocks of the last K3t launch
int tile_estimate_debug_cta_ns(unsigned long long* out1024);  // entry/exit ns of the first 512 CTAs
#endif
int tile_estimate_configure();  // one-time cudaFuncSetAttribute (dynamic smem); returns cudaError_t
// multi-GPU GLOBAL map (kernels_tile.cu)
void launch_shard_begin(ShardHeader* hdr, uint32_t seq, int world, int rank, uint32_t n_scan,
                        uint32_t back_weight, ShardSlice* slice_out, uint32_t* counters, cudaStream_t s,
                        LaunchCounter& lc);
// counters: the front counter block; words [kShardSlotBase, kShardSlotBase + world) are the record
// slots handed out so far in every owner's area
constexpr int kShardSlotBase = 16;
constexpr int kFrontCounterWords = 32;
void launch_shard_alloc(const TileBuffers& tb, uint32_t bps, int world, uint32_t* counters, cudaStream_t s,
                        LaunchCounter& lc);
void launch_shard_publish_front(const ShardFrontArgs& a, const uint32_t* counters, cudaStream_t s,
                                LaunchCounter& lc);
void launch_shard_gather(const ShardBackArgs& a, uint32_t* counters, cudaStream_t s, LaunchCounter& lc);
// light pass of K3t: one warp per bucket with <= 128 records; bigger buckets go to the heavy list
void launch_tile_estimate_light(const EstimateParams& p, const TileBuffers& tb, uint32_t* counters,
                                DeviceState* st_out, cudaStream_t s, LaunchCounter& lc);
void launch_tile_estimate_light_shard(const EstimateParams& p, const ShardBackArgs& a, uint32_t* counters,
                                      DeviceState* st_out, cudaStream_t s, LaunchCounter& lc);
void launch_tile_estimate_shard(const EstimateParams& p, const ShardBackArgs& a, uint32_t* counters,
                                DeviceState* st_out, const PublishArgs& pub, cudaStream_t s,
                                LaunchCounter& lc);
void launch_move_only(const DeviceState* st_in, DeviceState* st_out, double x, double y,
                      int clear_policy, const LayerTable& lt, uint32_t* moved_flag, cudaStream_t s,
                      LaunchCounter& lc);

// voxelGrid(ANY) + raycasting (kernels_raycast.cu)
void launch_voxel_keys(const float4* pm, uint32_t n, float inv_voxel, uint64_t* keys,
                       uint32_t* vals, cudaStream_t s, LaunchCounter& lc);
void launch_voxel_select(const uint64_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                         uint32_t* counters, uint32_t* out_sel, cudaStream_t s, LaunchCounter& lc);
// compact 32-bit voxel keys: box of the kept points predicted by the host from the crop filters
struct VoxelBox {
  int32_t x0, y0, z0;   // voxel coordinate of the box corner (after the reference's clamp)
  int32_t bx, by, bz;   // bits per axis; bx + by + bz <= 31
  uint32_t invalid_key; // 1 << (bx + by + bz): sorts after every valid key
};
void launch_voxel_keys32(const float4* pm, uint32_t n, float inv_voxel, const VoxelBox& box,
                         uint32_t* keys, uint32_t* vals, uint32_t* counters, cudaStream_t s,
                         LaunchCounter& lc);
void launch_voxel_select32(const uint32_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                           uint32_t invalid_key, uint32_t* counters, uint32_t* out_sel,
                           cudaStream_t s, LaunchCounter& lc);
// voxel representative + observed-evidence hits + list of the rays to trace, counting-sorted
// into (length, azimuth) bundles (sorted_keys == nullptr: every input point is a ray_scan point)
struct RaySortScratch {
  uint32_t* hist;      // [ray_sort_scratch_words()]: histogram (all zero between scans) + cursors
  float4* unsorted;    // [n] ray end points in discovery order, ordering key in .w
  float4* rays;        // [n] the sorted list the DDA kernel reads
};
size_t ray_sort_scratch_words();
void launch_voxel_select_rays32(const uint32_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                                uint32_t invalid_key, const RaycastParams& p, const DeviceState* st,
                                const float4* pts, uint32_t* counters, const RaySortScratch& rs,
                                cudaStream_t s, LaunchCounter& lc);
void launch_voxel_select_rays64(const uint64_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                                const RaycastParams& p, const DeviceState* st, const float4* pts,
                                uint32_t* counters, const RaySortScratch& rs, cudaStream_t s,
                                LaunchCounter& lc);
// scratch of the MSD voxel sort: `words` holds 5 row-sized arrays + 2 block-sized + 8 control words
// (voxel_rows_scratch_words(rows_cap)), all zero when first used and re-armed by the kernels
struct VoxelRowsScratch {
  uint32_t* words;
  uint32_t rows_cap;            // rows the arrays were sized for (a power of two, >= 1 << (by + bz))
  uint32_t* vkey;               // [n]
  unsigned long long* pairs;    // [n]
};
size_t voxel_rows_scratch_words(uint32_t rows_cap);
void launch_voxel_select_rays_msd(const float4* pm, uint32_t n, float inv_voxel, const VoxelBox& box,
                                  const VoxelRowsScratch& vs, const RaycastParams& p, const DeviceState* st,
                                  uint32_t* counters, const RaySortScratch& rs, cudaStream_t s, LaunchCounter& lc);
void launch_raycast_dda(const RaycastParams& p, const DeviceState* st, const float4* rays,
                        uint32_t n_max, uint32_t* counters, cudaStream_t s, LaunchCounter& lc);
void launch_raycast_resolve(const RaycastParams& p, const DeviceState* st, const LayerTable& lt,
                            const uint32_t* counters, size_t n_cells, cudaStream_t s,
                            LaunchCounter& lc);
int launch_median_filter(const float* src, float* dst, const DeviceState* st, int kernel_size,
                         int min_valid, cudaStream_t s, LaunchCounter& lc, int rows_local, int cols);
// returns 1 when radius / resolution exceeds the compiled neighbourhood (8 cells)
int launch_uncertainty_fusion(const float* upper_in, const float* lower_in, float* upper_out,
                              float* lower_out, const DeviceState* st, float radius, double res,
                              float spatial_sigma, float q_lower, float q_upper, int min_valid,
                              cudaStream_t s, LaunchCounter& lc, int rows_local, int cols);
int launch_feature_extraction(const float* elev, float* const out7[7], const DeviceState* st,
                              float radius, double res, int min_valid, float p_lo, float p_hi,
                              cudaStream_t s, LaunchCounter& lc, int rows_local, int cols);
// map -> PointCloud2 body.  phase 0: per-column counts + exclusive scan (+ total); phase 1:
// ordered write of `total` points of (3 + n_fields) floats
void launch_pack_pointcloud2(const float* elev, const float* const* fields, int n_fields, int rows,
                             int cols, int sub_r0, int sub_c0, int sub_rows, int sub_cols,
                             const DeviceState* st, uint32_t* col_count, uint32_t* col_offset,
                             uint32_t* total, float* out, int phase, cudaStream_t s,
                             LaunchCounter& lc);
void launch_inpaint_stripe(const float* src, float* dst, int rows_local, int cols, const float* above,
                           const float* below, int min_valid, cudaStream_t s, LaunchCounter& lc);
void launch_inpaint_iter(const float* src, float* dst, const DeviceState* st, int min_valid,
                         cudaStream_t s, LaunchCounter& lc, int rows_local, int cols);

// radix sorts (sort.cu; CUB DeviceRadixSort — stable, which the tie-breaks rely on)
size_t sort_pairs_u32_temp_bytes(uint32_t n, int end_bit);
cudaError_t sort_pairs_u32(void* temp, size_t temp_bytes, const uint32_t* kin, uint32_t* kout,
                           const uint32_t* vin, uint32_t* vout, uint32_t n, int end_bit,
                           cudaStream_t s, LaunchCounter& lc);
size_t sort_pairs_u64_temp_bytes(uint32_t n, int end_bit);
cudaError_t sort_pairs_u64(void* temp, size_t temp_bytes, const uint64_t* kin, uint64_t* kout,
                           const uint32_t* vin, uint32_t* vout, uint32_t n, int end_bit,
                           cudaStream_t s, LaunchCounter& lc);

}  // namespace fdem
