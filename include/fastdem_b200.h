/* fastdem_b200.h — C-ABI of libfastdem_b200.so
 *
 * The drop-in boundary for ONE path of Ikhyeon-Cho/FastDEM (reference @ 30d371e):
 * the per-scan point-cloud -> elevation-grid integration,
 *   fastdem::FastDEM::integrate            fastdem/include/fastdem/fastdem.hpp:115-120
 *   fastdem::ElevationMapping::update      fastdem/include/fastdem/mapping/elevation_mapping.hpp:41-42
 *   fastdem::applyRaycasting               fastdem/include/fastdem/postprocess/raycasting.hpp:49-51
 * executed as hand-written CUDA for sm_100a on a device-resident map.
 *
 * The reference has no FFI layer: its boundary is the C++ class API.  This header
 * is what a binding for that API talks to (plain C, opaque handles, pointers and
 * sizes, int status, no exceptions, no torch/Eigen types).  The reference-facing
 * C++ shell (fastdem_b200/host/fastdem/..., same class names as the reference) and
 * the Python mirror (fastdem_b200/api.py) are both thin layers over these calls;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns fdem_status (0 = OK) unless noted; on error
 *     fdem_last_error() holds a message (thread-local).  Nothing throws.
 *   - one caller at a time per handle, like the reference ("not thread-safe",
 *     fastdem.hpp:48-53).
 *   - matrices (layers) are float32, rows x cols, COLUMN-MAJOR (linear = col*rows+row),
 *     exactly nanogrid::Matrix = Eigen::MatrixXf (fastdem/src/io_npz.cpp:142-144).
 *   - 4x4 transforms are double[16] COLUMN-MAJOR = Eigen::Isometry3d::matrix().data().
 *   - point clouds: xyzw float32 N x 4 (w = 1; nanopcl::PointCloud::points(),
 *     nanopcl/core/point_cloud.hpp:37-38), optional intensity float32 N, optional rgb
 *     uint8 N x 3 (nanopcl::Color, core/types.hpp:46-52).  Pointers may be HOST or
 *     DEVICE memory (detected with cudaPointerGetAttributes); host buffers are staged
 *     through pinned memory on the mapper's stream.
 *   - the CUDA device is fixed at fdem_map_create(); there is NO CPU fallback: without a
 *     usable device every entry point fails with FDEM_ERR_CUDA.
 */
#ifndef FASTDEM_B200_H
#define FASTDEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define FDEM_ABI_VERSION 1

typedef int32_t fdem_status;
enum {
  FDEM_OK = 0,
  FDEM_ERR_INVALID_ARGUMENT = 1, /* bad pointer / size / enum; also nanopcl voxelGrid's
                                    std::invalid_argument (voxel_grid_impl.hpp:31-33) */
  FDEM_ERR_CUDA = 2,             /* any CUDA runtime error (message has the string) */
  FDEM_ERR_NO_LAYER = 3,         /* nanogrid::GridMap::get() on a missing layer */
  FDEM_ERR_OUT_OF_MEMORY = 4,
  FDEM_ERR_UNSUPPORTED = 5
};

/* fastdem::SensorType          fastdem/include/fastdem/config/sensor_model.hpp:10-14 */
enum { FDEM_SENSOR_CONSTANT = 0, FDEM_SENSOR_LIDAR = 1, FDEM_SENSOR_RGBD = 2 };
/* fastdem::MappingMode         fastdem/include/fastdem/config/mapping.hpp:8-11 */
enum { FDEM_MODE_LOCAL = 0, FDEM_MODE_GLOBAL = 1 };
/* fastdem::EstimationType      fastdem/include/fastdem/config/mapping.hpp:13-16 */
enum { FDEM_EST_KALMAN = 0, FDEM_EST_P2QUANTILE = 1 };
/* which layers GridMap::move() resets on the vacated rows/cols.  nanoGrid is not in the
 * reference tree, so this is an explicit, documented choice (DESIGN.md "nanoGrid"). */
enum { FDEM_MOVE_CLEAR_ALL_LAYERS = 0, FDEM_MOVE_CLEAR_BASIC_LAYERS = 1 };

/* fastdem::Config, flattened    fastdem/include/fastdem/config/fastdem.hpp:21-37
 * (+ config/mapping.hpp:20-47, config/sensor_model.hpp:18-37, config/postprocess.hpp:15-23).
 * Field defaults = the reference's struct defaults; fdem_config_default() fills them. */
typedef struct fdem_config {
  /* config::PointFilter */
  float z_min, z_max, range_min, range_max;
  /* config::SensorModel */
  int32_t sensor_type;
  float lidar_range_noise, lidar_angular_noise;
  float rgbd_normal_a, rgbd_normal_b, rgbd_normal_c, rgbd_lateral_factor;
  float constant_uncertainty;
  /* config::Mapping */
  int32_t mode;
  int32_t estimation_type;
  float kalman_min_variance, kalman_max_variance, kalman_process_noise;
  float p2_dn[5];
  int32_t p2_elevation_marker;
  float p2_max_sample_count;
  /* config::Raycasting */
  int32_t raycasting_enabled;
  float rc_height_conflict_threshold, rc_log_odds_observed, rc_log_odds_ghost, rc_log_odds_max,
      rc_clear_threshold;
  /* restated-nanoGrid switch (see enum above) */
  int32_t move_clear_policy;
} fdem_config;

/* What one integrate() did.  `integrated` is FastDEM::integrate's bool
 * (fastdem/src/fastdem.cpp:125-128,137-138,161). */
typedef struct fdem_scan_stats {
  int64_t n_input;  /* points passed in                                         */
  int64_t n_kept;   /* survived cropRange + cropZ (fastdem.cpp:175-176)          */
  int64_t n_cells;  /* cells touched by rasterize (elevation_mapping.cpp:41-92)  */
  int64_t n_voxels; /* ray_scan size after voxelGrid(ANY) (fastdem.cpp:156-157)  */
  int32_t integrated;
  int32_t voxel_box_violations; /* raycasting self-check: points outside the key box the
                                   host predicted from the crop filters; always 0 */
} fdem_scan_stats;

/* nanogrid::GridMap geometry accessors (getSize/getResolution/getLength/getPosition/
 * getStartIndex), as of the last completed call. */
typedef struct fdem_geometry {
  int32_t rows, cols;
  double resolution;
  double length[2];
  double position[2];
  int32_t start_index[2];
  /* row stripe held by this handle: logical rows [row_begin, row_end) of a rows x cols map.
   * [0, rows) for an unsharded map. */
  int32_t row_begin, row_end;
} fdem_geometry;

typedef struct fdem_map fdem_map;       /* fastdem::ElevationMap, device resident */
typedef struct fdem_mapper fdem_mapper; /* fastdem::FastDEM bound to one map       */

/* ── library ─────────────────────────────────────────────────────────────── */
int32_t fdem_abi_version(void);
const char* fdem_last_error(void);
const char* fdem_status_string(fdem_status s);
void fdem_config_default(fdem_config* cfg); /* = fastdem::Config{} */
/* detail::validate of the reference's YAML loader (fastdem/src/config_fastdem.cpp:128-260):
 * FDEM_ERR_INVALID_ARGUMENT where the reference throws (kalman min_variance >= max_variance,
 * unsorted P2 markers), values the reference warns about are clamped in place and counted in
 * *n_clamped (may be NULL).  A Config struct handed straight to FastDEM(map, cfg) is NOT
 * validated by the reference, and fdem_mapper_create does not validate it either. */
fdem_status fdem_config_validate(fdem_config* cfg, int32_t* n_clamped);

/* ── map: fastdem::ElevationMap  (fastdem/include/fastdem/elevation_map.hpp:65-177) ── */

/* ElevationMap(width, height, resolution, frame) + setGeometry (elevation_map.hpp:105-116):
 * size = round(length/resolution), basic layers elevation/elevation_min/elevation_max, all
 * NaN.  `device` = CUDA ordinal.  `stream` = a cudaStream_t the map's work is ordered on
 * (0 => the library creates its own non-blocking stream). */
fdem_status fdem_map_create(float width, float height, float resolution, int32_t device,
                            void* stream, fdem_map** out);
/* One row stripe [row_begin,row_end) of the same logical map, for the multi-GPU GLOBAL
 * configuration.  Only GLOBAL mapping is valid on a stripe. */
fdem_status fdem_map_create_stripe(float width, float height, float resolution,
                                   int32_t row_begin, int32_t row_end, int32_t device,
                                   void* stream, fdem_map** out);
/* Destroys the map.  FastDEM objects (fdem_mapper) still bound to it are detached first: their
 * calls then fail with FDEM_ERR_INVALID_ARGUMENT, fdem_mapper_destroy stays valid. */
fdem_status fdem_map_destroy(fdem_map* map);
/* ElevationMap::setGeometry(width, height, resolution) on an existing map, IN PLACE
 * (elevation_map.hpp:112-116: nanoGrid setGeometry + clearAll): every existing layer is resized
 * and reset to NaN, position and start index return to 0.  The handle — and every fdem_mapper
 * bound to it, like the reference's FastDEM holding an ElevationMap& — stays valid; this is
 * what io::loadNpz does to the map it restores into (fastdem/src/io_npz.cpp:440-).  Not valid
 * on a row stripe. */
fdem_status fdem_map_set_geometry(fdem_map* map, float width, float height, float resolution);

fdem_status fdem_map_get_geometry(fdem_map* map, fdem_geometry* out);
/* GridMap::setPosition / setStartIndex (used by snapshot/load, elevation_map.hpp:166-167) */
fdem_status fdem_map_set_position(fdem_map* map, double x, double y);
fdem_status fdem_map_set_start_index(fdem_map* map, int32_t row, int32_t col);
/* GridMap::move(position): circular-buffer shift, vacated rows/cols reset to NaN
 * (call site fastdem/src/elevation_mapping.cpp:111-113).  *moved = 1 if the start index changed. */
fdem_status fdem_map_move(fdem_map* map, double x, double y, int32_t clear_policy, int32_t* moved);
/* GridMap::isInside / getIndex / getPosition (host-side geometry math, no device work) */
fdem_status fdem_map_is_inside(fdem_map* map, double x, double y, int32_t* inside);
fdem_status fdem_map_get_index(fdem_map* map, double x, double y, int32_t* row, int32_t* col,
                               int32_t* inside);
fdem_status fdem_map_get_cell_position(fdem_map* map, int32_t row, int32_t col, double* x,
                                       double* y);

/* GridMap::exists / add(name, fill) / getLayers */
fdem_status fdem_map_layer_exists(fdem_map* map, const char* name, int32_t* exists);
fdem_status fdem_map_layer_add(fdem_map* map, const char* name, float fill);
fdem_status fdem_map_layer_count(fdem_map* map, int32_t* count);
fdem_status fdem_map_layer_name(fdem_map* map, int32_t i, char* buf, int32_t cap);
/* GridMap::get(layer) as a copy: dst/src are (row_end-row_begin) x cols float32 column-major,
 * HOST or DEVICE memory. */
fdem_status fdem_map_layer_download(fdem_map* map, const char* name, float* dst);
fdem_status fdem_map_layer_upload(fdem_map* map, const char* name, const float* src);
/* zero-copy view for torch / downstream kernels: device pointer to the layer slab */
fdem_status fdem_map_layer_device_ptr(fdem_map* map, const char* name, float** dptr);
/* GridMap::at(layer, index) read / write of a single cell (tests, clearAt fixtures) */
fdem_status fdem_map_cell_get(fdem_map* map, const char* name, int32_t row, int32_t col, float* v);
fdem_status fdem_map_cell_set(fdem_map* map, const char* name, int32_t row, int32_t col, float v);
/* GridMap::clear(layer) / clearAll() (FastDEM::reset, fastdem/src/fastdem.cpp:26) /
 * ElevationMap::clearAt (elevation_map.hpp:131-135) / ElevationMap::isEmpty (:123-125) */
fdem_status fdem_map_clear(fdem_map* map, const char* name);
fdem_status fdem_map_clear_all(fdem_map* map);
fdem_status fdem_map_clear_at(fdem_map* map, int32_t row, int32_t col);
fdem_status fdem_map_is_empty(fdem_map* map, int32_t* empty);
fdem_status fdem_map_sync(fdem_map* map); /* wait for all queued work on the map's stream */
void* fdem_map_stream(fdem_map* map);     /* the cudaStream_t work is ordered on */

/* ── mapper: fastdem::FastDEM  (fastdem/include/fastdem/fastdem.hpp:55-156) ──────── */

/* FastDEM(ElevationMap&, const Config&) (fastdem/src/fastdem.cpp:19-22): creates the
 * estimator layers (Kalman::ensureLayers kalman_estimation.hpp:64-82 or
 * P2Quantile::ensureLayers quantile_estimation.hpp:97-115) and `obstacle`
 * (elevation_mapping.cpp:38).  The map must outlive the mapper (FastDEM holds a reference). */
fdem_status fdem_mapper_create(fdem_map* map, const fdem_config* cfg, fdem_mapper** out);
fdem_status fdem_mapper_destroy(fdem_mapper* m);
/* the fluent setters (fastdem.cpp:28-66) collapsed into one call: re-creates the sensor
 * model / ElevationMapping exactly as setMappingMode/setEstimatorType/setSensorModel do
 * (missing estimator layers are added, existing data persists). */
fdem_status fdem_mapper_set_config(fdem_mapper* m, const fdem_config* cfg);
fdem_status fdem_mapper_get_config(fdem_mapper* m, fdem_config* out);

/* FastDEM::integrate(cloud, T_base_sensor, T_world_base) (fastdem.cpp:122-162), synchronous:
 * returns after the scan is in the map, *stats filled.  stats->integrated == 0 <=> the
 * reference returns false (empty input / everything filtered). */
fdem_status fdem_mapper_integrate(fdem_mapper* m, const float* xyzw, const float* intensity,
                                  const uint8_t* rgb, size_t n, const double T_base_sensor[16],
                                  const double T_world_base[16], fdem_scan_stats* stats);
/* Same work, queued on the map's stream without waiting; results of scan k are read with
 * fdem_mapper_wait().  Lets the caller keep the GPU fed (no host round trip per scan).
 * Device input buffers must stay valid until the matching wait. */
fdem_status fdem_mapper_integrate_async(fdem_mapper* m, const float* xyzw, const float* intensity,
                                        const uint8_t* rgb, size_t n,
                                        const double T_base_sensor[16],
                                        const double T_world_base[16]);
/* n_scans (1..16) consecutive integrate() calls in one go — for callers that have several scans at
 * hand (bag replay, several sensors per tick).  The result is exactly that of calling
 * fdem_mapper_integrate n_scans times; what changes is the schedule: the scans go to the device
 * as ONE graph in which scan k+1's transform / binning / partition kernels run beside scan k's
 * per-cell estimator (they touch only scan scratch and the chain of window geometries), so a
 * scan costs max(front, back) instead of front + back.  Arrays of n_scans entries; transforms
 * are n_scans x 16 doubles; intensity / rgb may be NULL (no such channel) — all scans of a batch
 * carry the same channels.  Device-resident inputs take the overlapped schedule; host inputs,
 * raycasting and the global-sort path fall back to scan-after-scan.  stats: n_scans entries, or
 * NULL to queue the batch without waiting (then see fdem_mapper_wait). */
fdem_status fdem_mapper_integrate_batch(fdem_mapper* m, int32_t n_scans, const float* const* xyzw,
                                        const float* const* intensity, const uint8_t* const* rgb,
                                        const size_t* num_points, const double* T_base_sensor,
                                        const double* T_world_base, fdem_scan_stats* stats);
/* per-scan statistics of the most recent batch (waits for it): what a queued batch
 * (stats == NULL) would have returned.  Every scan's counters reach pinned host memory at the end
 * of the batch, as they do scan by scan. */
fdem_status fdem_mapper_last_batch_stats(fdem_mapper* m, fdem_scan_stats* stats, int32_t n_scans);
/* blocks until every queued scan is done; *stats (optional) = the LAST scan's stats. */
fdem_status fdem_mapper_wait(fdem_mapper* m, fdem_scan_stats* stats);
/* Streaming form of the same call: submit() queues a scan and returns its ticket at once;
 * collect() waits for THAT scan only and returns its stats.  Host inputs are copied on a
 * separate stream into a double-buffered staging area, so `submit(k+1); collect(k);` overlaps
 * the host->device copy of scan k+1 with the kernels of scan k.  Input buffers must stay
 * valid until the scan's collect() returns; at most 8 scans may be outstanding. */
fdem_status fdem_mapper_submit(fdem_mapper* m, const float* xyzw, const float* intensity,
                               const uint8_t* rgb, size_t n, const double T_base_sensor[16],
                               const double T_world_base[16], uint64_t* ticket);
fdem_status fdem_mapper_collect(fdem_mapper* m, uint64_t ticket, fdem_scan_stats* stats);

/* ElevationMapping::update(cloud, robot_position) (elevation_mapping.cpp:110-125): points
 * already in the map frame; var_z optional = cloud.covariance(i)(2,2) (null => cloud has no
 * covariance channel => 0 => Kalman uses max_variance, kalman_estimation.hpp:112-113). */
fdem_status fdem_mapper_update(fdem_mapper* m, const float* xyzw, const float* var_z,
                               const float* intensity, const uint8_t* rgb, size_t n,
                               double robot_x, double robot_y, fdem_scan_stats* stats);

/* integrate() for a caller-provided SensorModel subclass (FastDEM::setSensorModel(unique_ptr),
 * fastdem.hpp:80): cov9 = N x 9 float32 column-major sensor-frame covariances computed by the
 * caller's computeCovariances() (sensors/sensor_model.hpp:76-85). */
fdem_status fdem_mapper_integrate_with_cov(fdem_mapper* m, const float* xyzw, const float* cov9,
                                           const float* intensity, const uint8_t* rgb, size_t n,
                                           const double T_base_sensor[16],
                                           const double T_world_base[16], fdem_scan_stats* stats);

/* The caller-side data format: a sensor_msgs/PointCloud2 message body handed over as is
 * (what ros2/src/fastdem_ros_node.cpp:178-198 receives and nanopcl::from(msg) unpacks on the
 * CPU, fastdem/lib/nanoPCL/include/nanopcl/bridge/ros/impl.hpp:180-270).  The unpacking —
 * x/y/z at their field offsets, points with a non-finite coordinate dropped, intensity
 * converted from its PointField datatype, packed 0x00RRGGBB colour — happens inside the
 * first kernel, so the message crosses PCIe once, as one contiguous copy.
 * Offsets as FieldOffsets::parse finds them (-1 = field absent).  4-byte fields must be
 * 4-byte aligned within the point (they are in every message ROS drivers emit). */
typedef struct fdem_pointcloud2_layout {
  uint32_t point_step;     /* msg.point_step                                                */
  int32_t off_x, off_y, off_z;
  int32_t off_intensity;   /* "intensity"                                                   */
  int32_t intensity_type;  /* PointField datatype: 2 UINT8, 4 UINT16, 7 FLOAT32, 8 FLOAT64   */
  int32_t off_rgb;         /* "rgb" / "rgba"                                                */
} fdem_pointcloud2_layout;
/* integrate(nanopcl::from(msg), T_base_sensor, T_world_base); data = msg.data (HOST or DEVICE),
 * num_points = msg.width * msg.height.  stats->n_input = points with finite coordinates. */
fdem_status fdem_mapper_integrate_pointcloud2(fdem_mapper* m, const uint8_t* data,
                                              size_t num_points,
                                              const fdem_pointcloud2_layout* layout,
                                              const double T_base_sensor[16],
                                              const double T_world_base[16],
                                              fdem_scan_stats* stats);
/* streaming form (see fdem_mapper_submit): queue the message, collect its stats by ticket; the
 * copy of message k+1 overlaps the kernels of message k.  A PointCloud2 point is usually
 * smaller than the Vector4f + intensity the reference unpacks it into (x, y, z, intensity =
 * 16 bytes instead of 20), so this is also the cheapest way across PCIe. */
fdem_status fdem_mapper_submit_pointcloud2(fdem_mapper* m, const uint8_t* data, size_t num_points,
                                           const fdem_pointcloud2_layout* layout,
                                           const double T_base_sensor[16],
                                           const double T_world_base[16], uint64_t* ticket);

/* FastDEM::onScanPreprocessed payload (fastdem.cpp:139-141): the preprocessed cloud of the
 * LAST integrate, compacted in input order.  Buffers are HOST memory sized for n_kept:
 * xyzw [n_kept*4], cov9 [n_kept*9] (optional), src_index [n_kept] (optional). */
fdem_status fdem_mapper_last_preprocessed(fdem_mapper* m, float* xyzw, float* cov9,
                                          int32_t* src_index, int64_t* n_kept);
/* FastDEM::onScanRasterized payload (fastdem.cpp:148-150, 200-214): one point per touched
 * cell (cell centre x,y; z = min_z) of the LAST integrate.  xyz HOST [n_cells*3]; order is by
 * ascending buffer linear index (the reference's order is hash order, i.e. unspecified). */
fdem_status fdem_mapper_last_rasterized(fdem_mapper* m, float* xyz, int64_t* n_cells);

/* ── stage-level entry points ─────────────────────────────────────────────── */

/* fastdem::applyRaycasting(map, scan, sensor_origin, cfg) (fastdem/src/raycasting.cpp:218-249);
 * scan in map frame, HOST or DEVICE. */
fdem_status fdem_raycast(fdem_map* map, const float* xyzw, size_t n, const float sensor_origin[3],
                         const fdem_config* cfg);
/* nanopcl::filters::voxelGrid(cloud, voxel_size, VoxelMode::ANY) (voxel_grid_impl.hpp:30-236):
 * out_index HOST [<= n] = selected source indices in voxel-key order; ties resolved as the
 * oracle defines (sort on (key, index)). */
fdem_status fdem_voxel_grid_any(fdem_map* map, const float* xyzw, size_t n, float voxel_size,
                                uint32_t* out_index, int64_t* n_voxels);
/* One Jacobi sweep of applyInpainting (fastdem/src/inpainting.cpp:41-62) on a ROW STRIPE of a
 * GLOBAL map, in place on `layer` (snapshot in, layer out).  The logical rows just above / below
 * the stripe live on the neighbouring ranks: they are passed as DEVICE pointers to `cols` floats
 * each (NULL at the map border) — the 1-row halo the sharded driver exchanges (NCCL send / recv on
 * the map's stream) between sweeps.  Asynchronous on the map's stream. */
fdem_status fdem_inpaint_stripe_sweep(fdem_map* map, const char* layer, const float* halo_above,
                                      const float* halo_below, int32_t min_valid_neighbors);
/* fastdem::applyInpainting(map, max_iterations, min_valid_neighbors, inplace)
 * (fastdem/src/inpainting.cpp:21-67) */
fdem_status fdem_inpaint(fdem_map* map, int32_t max_iterations, int32_t min_valid_neighbors,
                         int32_t inplace);

/* fastdem::applySpatialSmoothing(map, layer, kernel_size, min_valid_neighbors)
 * (fastdem/include/fastdem/postprocess/spatial_smoothing.hpp:38-67): in-place median filter;
 * a missing layer is a no-op like in the reference.  kernel_size in {1, 3, 5, 7}. */
fdem_status fdem_spatial_smoothing(fdem_map* map, const char* layer_name, int32_t kernel_size,
                                   int32_t min_valid_neighbors);

/* fastdem::applyUncertaintyFusion(map, config::UncertaintyFusion)
 * (fastdem/src/uncertainty_fusion.cpp:103-186; config/postprocess.hpp:33-40): bilateral-weighted
 * quantiles of the neighbours' lower / upper bounds, written back to upper_bound / lower_bound.
 * Missing bound layers are a no-op like in the reference (warn + return).  The neighbourhood
 * (search_radius / resolution) may span at most 8 cells: FDEM_ERR_UNSUPPORTED beyond. */
fdem_status fdem_uncertainty_fusion(fdem_map* map, float search_radius, float spatial_sigma,
                                    float quantile_lower, float quantile_upper,
                                    int32_t min_valid_neighbors);

/* fastdem::applyFeatureExtraction(map, analysis_radius, min_valid_neighbors,
 * step_lower_percentile, step_upper_percentile) (fastdem/src/feature_extraction.cpp:28-118):
 * local PCA of the elevation patch -> layers step, slope, roughness, curvature, _normal_x,
 * _normal_y, _normal_z (added NaN-filled when missing).  No elevation layer = no-op. */
fdem_status fdem_feature_extraction(fdem_map* map, float analysis_radius,
                                    int32_t min_valid_neighbors, float step_lower_percentile,
                                    float step_upper_percentile);

/* fastdem::ros::toPointCloud2Impl(map, stamp, elevation_layer, sub_start, sub_size)
 * (fastdem/include/fastdem/bridge/ros/impl.hpp:29-174): the map as a sensor_msgs/PointCloud2
 * body, packed on the device.  One point per cell of the sub-region whose `elevation_layer`
 * value is finite; fields (all FLOAT32, 4 bytes, in this order): x, y, z, every layer whose
 * name does not start with '_' except the elevation layer and "color", then "rgb" (the packed
 * colour bits) if the map has a colour layer.  Points are ordered columns outer / rows inner
 * from (sub_row, sub_col), wrapping around the circular buffer.  sub_rows = sub_cols = -1
 * selects the full map (start = the buffer start index), the reference's convenience overload
 * (:170-174).  The packed body stays on the device until the next pack; read it with
 * fdem_map_pointcloud2_data (dst: host or device, width * point_step bytes; device_ptr:
 * optional zero-copy view) and its field names with fdem_map_pointcloud2_field. */
fdem_status fdem_map_pack_pointcloud2(fdem_map* map, const char* elevation_layer, int32_t sub_row,
                                      int32_t sub_col, int32_t sub_rows, int32_t sub_cols,
                                      uint32_t* width, uint32_t* point_step, int32_t* n_fields);
fdem_status fdem_map_pointcloud2_field(fdem_map* map, int32_t i, char* buf, int32_t cap,
                                       uint32_t* offset);
fdem_status fdem_map_pointcloud2_data(fdem_map* map, uint8_t* dst, const uint8_t** device_ptr);

/* ── multi-GPU plumbing (no counterpart in the single-process reference) ──
 * One process per GPU.  A scan that every row stripe needs lives ONCE, in the ingest rank's
 * HBM; the other ranks map that buffer (CUDA IPC, peer access over NVLink / NVSwitch) and pass
 * the mapped pointer to fdem_mapper_integrate*: K1 and the scatter kernel then read the points
 * straight out of the peer's memory while they bin them — the "broadcast" is fused into the
 * first kernel that needs the data instead of being a collective of its own. */
typedef struct fdem_ipc_handle {
  uint8_t bytes[64]; /* cudaIpcMemHandle_t */
  uint64_t size;
} fdem_ipc_handle;
/* Bind the calling host thread to the CPUs local to the device's NUMA node (sysfs
 * local_cpulist); *n_cpus (may be NULL) = CPUs in the set.  Call it before allocating the pinned
 * scan buffers: one process per GPU on a multi-socket box otherwise pins everything on node 0.
 * FDEM_ERR_UNSUPPORTED when the topology is not exposed (containers without sysfs). */
fdem_status fdem_bind_thread_to_device(int32_t device, int32_t* n_cpus);
fdem_status fdem_device_alloc(int32_t device, size_t bytes, void** ptr);
fdem_status fdem_device_free(int32_t device, void* ptr);
/* `ptr` must come from fdem_device_alloc (IPC handles name whole allocations) */
fdem_status fdem_ipc_export(int32_t device, const void* ptr, size_t bytes, fdem_ipc_handle* out);
/* maps the exporting process's buffer into this process; enables peer access on first use */
fdem_status fdem_ipc_import(int32_t device, const fdem_ipc_handle* h, void** ptr);
fdem_status fdem_ipc_close(int32_t device, void* ptr);

/* ── multi-GPU GLOBAL map with compute that scales ────────────────────────
 * One logical map, row-striped over `world` ranks, one process per GPU.  The reference has no
 * distributed code; this is the B200-side answer to BASELINE.json's config 5.  Each rank holds
 * its stripe (fdem_map_create_stripe with the rows fastdem_b200/sharded.py's stripe_bounds gives
 * it), a FastDEM on it (GLOBAL mode) and one fdem_shard: the exchange arena every other rank maps
 * over CUDA IPC.  A scan is integrated in two halves per rank:
 *   front (the mapper's side stream): preprocessScan + binning of this rank's SLICE of the scan's
 *          points for EVERY stripe; each pre-reduced 32-byte record is stored straight into its
 *          owner's arena (peer memory: NVLink stores).  The slice is decided on the device from
 *          the stripes' loads two scans earlier — the same numbers on every rank, so the same
 *          split — such that a rank whose stripe owns most of the touched cells (most of the
 *          back half) bins few points or none;
 *   back (the map's stream): for every non-empty bucket of this rank's OWN stripe, the per-cell
 *          estimator over the records all sources pushed — local memory only.
 * Device-side ready / consumed flags (system-scope release / acquire on words of the arenas)
 * order the halves across ranks: no collective, no host round trip.  Results are identical to
 * one unsharded map (tests/test_gpu_shard.py; tools/shard_parity.py on real GPUs).
 * Setup: create on every rank -> export -> exchange the handles out of band (e.g.
 * torch.distributed.all_gather_object) -> connect. */
typedef struct fdem_shard fdem_shard;
fdem_status fdem_shard_create(fdem_mapper* mapper, int32_t rank, int32_t world, size_t max_points,
                              fdem_shard** out);
fdem_status fdem_shard_destroy(fdem_shard* shard);
fdem_status fdem_shard_export(fdem_shard* shard, fdem_ipc_handle* out);
/* handles[world], indexed by rank (handles[rank] is ignored) */
fdem_status fdem_shard_connect(fdem_shard* shard, const fdem_ipc_handle* handles);
/* FastDEM::integrate(cloud, T_base_sensor, T_world_base) on the striped map, asynchronous.
 * EVERY rank calls it for every scan, in the same order, with the same scan: xyzw / intensity /
 * rgb address the WHOLE scan (n points) in device memory this rank can read; the scan must be
 * complete there when the call is made (the halves run on two streams and are not ordered behind
 * earlier work of the map's stream) and stay untouched until fdem_shard_wait has returned.
 * Any rank may be handed most of a scan: give every rank a copy in its own HBM when NVLink reads
 * of the scan would bound its K1 (sharded.PeerScanRing(replicated=True)). */
fdem_status fdem_shard_integrate(fdem_shard* shard, const float* xyzw, const float* intensity,
                                 const uint8_t* rgb, size_t n, const double T_base_sensor[16],
                                 const double T_world_base[16]);
/* The slice split of the front half, on the host (pure arithmetic, no GPU needed): given the
 * cells each stripe's owner touched (loads[world]) it returns, for every rank, the first point
 * and the number of points of an n_points scan that rank bins — slices tile [0, n_points) in rank
 * order, boundaries are multiples of 32, ranks with busy stripes get fewer points.
 * back_weight_q8 = cost of a whole back half in units of a whole front half, x256 (0 = default). */
fdem_status fdem_shard_slice_plan(const uint32_t* loads, int32_t world, uint32_t n_points,
                                  uint32_t back_weight_q8, uint32_t* begins, uint32_t* counts);
/* waits for this rank's queued scans; stats of the newest as THIS rank saw it: n_kept = kept
 * points of its slice, n_cells = touched cells of its stripe (sum over ranks = scan totals),
 * n_input = the whole scan */
fdem_status fdem_shard_wait(fdem_shard* shard, fdem_scan_stats* stats);

/* ── instrumentation ──────────────────────────────────────────────────────── */
/* pipeline stages of one scan, in stream order */
enum {
  FDEM_STAGE_H2D = 0,        /* host -> device staging of the cloud (0 when inputs are device) */
  FDEM_STAGE_PREPROCESS = 1, /* K1 preprocess_bin_kernel                                       */
  FDEM_STAGE_COMMIT = 2,     /* K2 commit_move_clear_kernel                                    */
  FDEM_STAGE_SORT = 3,       /* sort by cell                                                   */
  FDEM_STAGE_ESTIMATE = 4,   /* K3 segreduce_estimate_kernel                                   */
  FDEM_STAGE_RAYCAST = 5,    /* voxelGrid(ANY) + raycasting (0 when disabled)                  */
  FDEM_STAGE_COUNT = 6
};
/* When enabled, every scan brackets its stages with CUDA events on the map's stream;
 * fdem_mapper_stage_times() returns the accumulated device milliseconds per stage and the
 * number of scans accumulated, and resets the accumulators.  Implies a stream sync. */
fdem_status fdem_mapper_set_stage_timing(fdem_mapper* m, int32_t enabled);
fdem_status fdem_mapper_stage_times(fdem_mapper* m, double ms[FDEM_STAGE_COUNT], int64_t* scans);
/* Which sort-by-cell implementation the mapper uses (results are identical):
 *   TILE   (default) 2-level sort: bucket partition + per-bucket shared-memory sort and
 *          warp-segmented reduce, TMA-staged (kernels_tile.cu)
 *   GLOBAL one CUB radix sort of the whole scan + warp-segmented reduce (kernels.cu) */
enum { FDEM_CELL_SORT_TILE = 0, FDEM_CELL_SORT_GLOBAL = 1 };
fdem_status fdem_mapper_set_cell_sort(fdem_mapper* m, int32_t mode);
/* Which sort voxelGrid(ANY) uses on the raycasting path (results are identical):
 *   MSD     2-level sort written for it: rows (z, y) by histogram + scan + scatter, then
 *           (x, index) inside each row on chip — warp registers / shared memory (kernels_raycast.cu);
 *           no library launches, slower on dense scans
 *   LIBRARY (default) cub::DeviceRadixSort on 32-bit box-relative keys (on the reference's 63-bit
 *           keys when the crop filters give no usable bound on the voxel box) */
enum { FDEM_VOXEL_SORT_MSD = 0, FDEM_VOXEL_SORT_LIBRARY = 1 };
fdem_status fdem_mapper_set_voxel_sort(fdem_mapper* m, int32_t mode);
/* kernels launched through CUB (radix-sort passes) since the map was created */
fdem_status fdem_mapper_library_launch_count(fdem_mapper* m, int64_t* launches);
/* number of kernels THIS library launched since the handle was created (bench.py's
 * gpu_launches) */
fdem_status fdem_mapper_launch_count(fdem_mapper* m, int64_t* launches);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FASTDEM_B200_H */
