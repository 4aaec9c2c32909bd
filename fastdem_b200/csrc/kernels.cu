// kernels.cu — the integrate() hot path as sm_100a CUDA kernels.
//
//   K1 preprocess_bin_kernel      sensor covariance + SE(3) x2 + range/height crop +
//                                 sigma_z^2 = (R S R^T)(2,2) + FP64 (x,y) -> cell key
//                                 (+ per-bucket point histogram on the tile path)
//   K2 commit_move_clear_kernel   LOCAL-mode circular-buffer move (vacated stripes -> NaN),
//                                 reset of last scan's obstacle cells, scan-state commit
//                                 (+ bucket segment allocation on the tile path)
//   K3 segreduce_estimate_kernel  global-sort path: warp-segmented reduce over the
//                                 CUB-sorted stream + one Kalman / P2 step per touched cell
//   (tile path: kernels_tile.cu;  global sort: sort.cu)
//
// Compiled with -fmad=false: every float/double expression below is evaluated exactly as
// written (no FMA contraction), in the operation order the CPU oracle fixes, so cell
// indices are bit-identical and heights/variances agree to the last bit in practice.
// Reference file:line citations are relative to /root/reference/.
#include <float.h>
#include <cstdlib>
#include <math.h>

#include "device_types.h"
#include "estimator.cuh"

namespace fdem {

namespace {

constexpr int kBlock = 256;

// ───────────────────────────── utility kernels ───────────────────────────────

__global__ void __launch_bounds__(kBlock) fill_kernel(float* __restrict__ dst, size_t n, float v) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}

__global__ void __launch_bounds__(kBlock) fill_u32_kernel(uint32_t* __restrict__ dst, size_t n,
                                                          uint32_t v) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}

// ElevationMap::isEmpty (elevation_map.hpp:123-125): flag = 1 if any cell is not NaN
__global__ void __launch_bounds__(kBlock) any_not_nan_kernel(const float* __restrict__ src,
                                                             size_t n, uint32_t* flag) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  bool any = false;
  for (; i < n; i += stride) any |= !isnan(src[i]);
  if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1u;
}

// ElevationMap::clearAt (elevation_map.hpp:131-135)
__global__ void clear_cell_kernel(LayerTable lt, int64_t lin) {
  const int l = threadIdx.x;
  if (l < lt.count) lt.ptr[l][lin] = nan_f32();
}

// ───────────────────────────── K1: preprocess + bin ──────────────────────────

// Matrix4f * Vector4f in the order r = c0*x; r = c1*y + r; r = c2*z + r; r = c3*w + r
// (nanopcl/core/transform.hpp:26-28; SURVEY.md §8a a6 "FP order")
__device__ __forceinline__ float4 transform_point(const float* __restrict__ T, float4 p) {
  float4 r;
  float acc;
  acc = T[0] * p.x;  acc = T[4] * p.y + acc;  acc = T[8] * p.z + acc;   acc = T[12] * p.w + acc;  r.x = acc;
  acc = T[1] * p.x;  acc = T[5] * p.y + acc;  acc = T[9] * p.z + acc;   acc = T[13] * p.w + acc;  r.y = acc;
  acc = T[2] * p.x;  acc = T[6] * p.y + acc;  acc = T[10] * p.z + acc;  acc = T[14] * p.w + acc;  r.z = acc;
  acc = T[3] * p.x;  acc = T[7] * p.y + acc;  acc = T[11] * p.z + acc;  acc = T[15] * p.w + acc;  r.w = acc;
  return r;
}

__device__ __forceinline__ float sqnorm3(float x, float y, float z) { return x * x + (y * y + z * z); }
__device__ __forceinline__ float dot3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

// sensor-frame covariance S (column-major 3x3), one of the three built-in models
__device__ __forceinline__ void sensor_covariance(const PreprocessParams& p, float4 q, float* S) {
#pragma unroll
  for (int k = 0; k < 9; ++k) S[k] = 0.0f;
  if (p.sensor_type == 1) {
    // LiDARSensorModel::computeCovariance (sensors/lidar_model.hpp:64-89)
    const float dist_sq = sqnorm3(q.x, q.y, q.z);
    if (dist_sq < 1e-6f) {
      S[0] = S[4] = S[8] = 0.01f;
      return;
    }
    const float distance = sqrtf(dist_sq);
    const float dir[3] = {q.x / distance, q.y / distance, q.z / distance};
    const float var_radial = fmaxf(p.lidar_range_noise * p.lidar_range_noise, 1e-6f);
    const float da = distance * p.lidar_angular_noise;
    const float var_lateral = fmaxf(da * da, 1e-6f);
    const float s = var_radial - var_lateral;
    const float sd[3] = {s * dir[0], s * dir[1], s * dir[2]};
    S[0] = S[4] = S[8] = var_lateral;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) S[j * 3 + i] = S[j * 3 + i] + dir[j] * sd[i];
  } else if (p.sensor_type == 2) {
    // RGBDSensorModel::computeCovariance (sensors/rgbd_model.hpp:82-101)
    const float depth = q.z;
    if (depth <= 0.0f) {
      S[0] = S[4] = S[8] = 0.01f;
      return;
    }
    const float diff = depth - p.rgbd_c;
    const float sigma_norm = p.rgbd_a + p.rgbd_b * diff * diff;
    const float sigma_lat = p.rgbd_k * depth;
    S[0] = S[4] = sigma_lat * sigma_lat;
    S[8] = sigma_norm * sigma_norm;
  } else {
    // ConstantUncertaintyModel (sensors/sensor_model.hpp:87-93)
    S[0] = S[4] = S[8] = p.constant_variance;
  }
}

// element (2,2) of R*S*R^T evaluated as the oracle does: tmp = R*S first, then tmp*R^T,
// 3-term dots as a0 + (a1 + a2)   (fastdem/src/fastdem.cpp:184-187; only cov(2,2) is
// consumed downstream, elevation_mapping.cpp:58-60)
__device__ __forceinline__ float rotated_var_z(const float* __restrict__ R, const float* S) {
  const float r0 = R[2], r1 = R[5], r2 = R[8];  // row 2 of column-major R
  const float t0 = dot3(r0 * S[0], r1 * S[1], r2 * S[2]);
  const float t1 = dot3(r0 * S[3], r1 * S[4], r2 * S[5]);
  const float t2 = dot3(r0 * S[6], r1 * S[7], r2 * S[8]);
  return dot3(t0 * r0, t1 * r1, t2 * r2);
}

__global__ void __launch_bounds__(kBlock)
preprocess_bin_kernel(const __grid_constant__ PreprocessParams p,
                      const DeviceState* __restrict__ st_in, uint32_t* __restrict__ counters,
                      float4* __restrict__ pm, uint32_t* __restrict__ keys,
                      uint32_t* __restrict__ vals) {
  // Geometry this scan bins against: LOCAL mode moves the window to the robot first
  // (elevation_mapping.cpp:111-113).  The move is committed by K2 only if >= 1 point
  // survives the filters (fastdem.cpp:137-138); if none does there is nothing to bin,
  // so binning against the prospective geometry is always right.
  __shared__ GridGeom sg;
  __shared__ uint32_t s_kept, s_inside;
  pdl_launch_dependents();  // K2 may start launching; it waits for this grid at its pdl_wait()
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  // multi-GPU front half: the slice was decided on the device (shard_begin_kernel); the grid
  // covers the whole scan and the CTAs beyond the slice leave at once
  uint32_t n_pts = p.n, in_base = 0;
  if (p.slice) {
    n_pts = p.slice->count;
    in_base = p.slice->begin;
    if (blockIdx.x * blockDim.x >= n_pts) return;
  }
  // the point load does not depend on the geometry: put it in flight first
  float4 q = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  bool finite = true;
  if (i < n_pts) {
    if (p.raw) {
      // nanopcl::from(msg): x/y/z floats at their field offsets, w = 1; non-finite points are
      // skipped before anything else sees them (bridge/ros/impl.hpp:237-244)
      const uint8_t* pt = p.raw + static_cast<size_t>(i) * p.point_step;
      if (p.raw_vec16) {
        // the common driver layout — x, y, z and one 4-byte field in a 16-byte point: ONE
        // 128-bit load per point, like the xyzw path
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.raw) + i);
        q.x = v.x; q.y = v.y; q.z = v.z;
        finite = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
        q.w = 1.0f;
        if (p.out_intensity && p.intensity_type == 7) {
          p.out_intensity[i] = v.w;
        } else if (p.out_rgb) {
          const uint32_t rgb = __float_as_uint(v.w);
          p.out_rgb[static_cast<size_t>(i) * 3 + 0] = static_cast<uint8_t>((rgb >> 16) & 0xFF);
          p.out_rgb[static_cast<size_t>(i) * 3 + 1] = static_cast<uint8_t>((rgb >> 8) & 0xFF);
          p.out_rgb[static_cast<size_t>(i) * 3 + 2] = static_cast<uint8_t>(rgb & 0xFF);
        }
      } else {
      q.x = __ldg(reinterpret_cast<const float*>(pt + p.off_x));
      q.y = __ldg(reinterpret_cast<const float*>(pt + p.off_y));
      q.z = __ldg(reinterpret_cast<const float*>(pt + p.off_z));
      q.w = 1.0f;
      finite = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
      if (p.out_intensity) {
        // readIntensity (bridge/ros/impl.hpp:106-121)
        const uint8_t* f = pt + p.off_intensity;
        float v = 0.0f;
        switch (p.intensity_type) {
          case 2: v = static_cast<float>(__ldg(f)); break;
          case 4: v = static_cast<float>(__ldg(reinterpret_cast<const uint16_t*>(f))); break;
          case 7: v = __ldg(reinterpret_cast<const float*>(f)); break;
          case 8: {
            const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(f));
            const uint32_t hi = __ldg(reinterpret_cast<const uint32_t*>(f) + 1);
            v = static_cast<float>(__hiloint2double(static_cast<int>(hi), static_cast<int>(lo)));
            break;
          }
          default: break;
        }
        p.out_intensity[i] = v;
      }
      if (p.out_rgb) {
        // readRgb (bridge/ros/impl.hpp:170-177): packed 0x00RRGGBB in a 4-byte field
        const uint32_t rgb = __ldg(reinterpret_cast<const uint32_t*>(pt + p.off_rgb));
        p.out_rgb[static_cast<size_t>(i) * 3 + 0] = static_cast<uint8_t>((rgb >> 16) & 0xFF);
        p.out_rgb[static_cast<size_t>(i) * 3 + 1] = static_cast<uint8_t>((rgb >> 8) & 0xFF);
        p.out_rgb[static_cast<size_t>(i) * 3 + 2] = static_cast<uint8_t>(rgb & 0xFF);
      }
      }
    } else {
      q = __ldg(&p.xyzw[in_base + i]);
    }
  }
  if (threadIdx.x == 0) {
    GridGeom g = st_in->geom;
    if (p.local_mode) {
      MoveResult mr;
      g = geom_move(g, p.robot_x, p.robot_y, mr);
    }
    sg = g;
    s_kept = 0;
    s_inside = 0;
  }
  __syncthreads();

  bool kept = false, inside = false;
  uint32_t key = p.invalid_key;
  if (i < n_pts) {
    float var_z = 0.0f;
    if (p.input_frame == INPUT_SENSOR_FRAME) {
      // preprocessScan (fastdem/src/fastdem.cpp:164-190)
      float S[9];
      if (p.cov9) {
#pragma unroll
        for (int k = 0; k < 9; ++k) S[k] = __ldg(&p.cov9[static_cast<size_t>(in_base + i) * 9 + k]);
      } else {
        sensor_covariance(p, q, S);
      }
      q = transform_point(p.T1, q);                       // sensor -> base
      const float d2 = sqnorm3(q.x, q.y, q.z);            // cropRange, base frame
      kept = finite && (d2 >= p.range_min_sq && d2 <= p.range_max_sq) &&
             (q.z >= p.z_min && q.z <= p.z_max);          // cropZ, base frame
      if (kept) {
        q = transform_point(p.T2, q);                     // base -> map
        var_z = rotated_var_z(p.R, S);
      }
    } else {
      // ElevationMapping::update seam: points already in the map frame
      kept = true;
      if (p.var_z) var_z = __ldg(&p.var_z[i]);
    }

    if (kept) {
      int32_t row, col;
      if (geom_get_index(sg, static_cast<double>(q.x), static_cast<double>(q.y), row, col)) {
        if (p.shard.world > 1) {
          // multi-GPU front half: this rank bins its slice of the scan for EVERY stripe
          // (owner-major keys; stripes exist only for GLOBAL maps, buffer row == logical row)
          int32_t rb, rl;
          const int32_t d = shard_of_row(p.shard, row, rb, rl);
          key = static_cast<uint32_t>(d) * p.shard.stride +
                static_cast<uint32_t>(col) * static_cast<uint32_t>(rl) + static_cast<uint32_t>(row - rb);
          inside = true;
        } else {
          const int64_t lin = geom_linear(sg, row, col);
          if (lin >= 0) {
            key = static_cast<uint32_t>(lin);
            inside = true;
          }
        }
      }
      pm[i] = make_float4(q.x, q.y, q.z, var_z);
    } else {
      pm[i] = make_float4(nan_f32(), nan_f32(), nan_f32(), 0.0f);  // dropped by the crop filters
    }
    keys[i] = key;
    if (p.write_vals) vals[i] = i;
  }

  if (p.raw) {
    const uint32_t fin_m = __ballot_sync(0xffffffffu, finite && i < n_pts);
    if (lane == 0 && fin_m) atomicAdd(&counters[CNT_FINITE], __popc(fin_m));
  }
  const uint32_t kept_m = __ballot_sync(0xffffffffu, kept);
  const uint32_t inside_m = __ballot_sync(0xffffffffu, inside);

  // tile path, L1 histogram: points per bucket of 2^kBucketBits consecutive cell keys.
  // One atomic per (warp, bucket) — scan-sized scratch, never map state.
  if (p.bucket_count && inside) {
    const uint32_t bucket = key >> p.bucket_bits;
    const uint32_t peers = __match_any_sync(inside_m, bucket);
    if (lane == __ffs(peers) - 1) atomicAdd(&p.bucket_count[bucket], __popc(peers));
  }

  // block-level counts -> one atomic per block per counter (scan statistics)
  if (lane == 0) {
    if (kept_m) atomicAdd(&s_kept, __popc(kept_m));
    if (inside_m) atomicAdd(&s_inside, __popc(inside_m));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_kept) atomicAdd(&counters[CNT_KEPT], s_kept);
    if (s_inside) atomicAdd(&counters[CNT_INSIDE], s_inside);
  }
}

// ───────────────────────────── K2: commit / move / clear ─────────────────────

// updateObstacle's map_.clear(obstacle) (elevation_mapping.cpp:146).  Only cells the last
// observing scan touched can hold a value, so only those are reset — unless the caller wrote
// the layer behind the mapper's back (SF_OBSTACLE_DIRTY): then the whole layer is cleared, as
// the reference does.  Runs only when this scan has observations (update() returns early
// otherwise, :116-117).
__device__ __forceinline__ void reset_obstacle(float* __restrict__ obstacle,
                                               const uint32_t* __restrict__ touched_keys,
                                               uint32_t n_prev, uint32_t invalid_key, bool dirty,
                                               size_t cells, size_t tid, size_t nthreads) {
  if (dirty) {
    for (size_t c = tid; c < cells; c += nthreads) obstacle[c] = nan_f32();
    return;
  }
  for (size_t j = tid; j < n_prev; j += nthreads) {
    const uint32_t k = touched_keys[j];
    if (k != invalid_key) obstacle[k] = nan_f32();
  }
}

// vacated rows / columns (or the whole map) -> NaN on the selected layers.  The
// (layer, cell) space is flattened so the stores spread evenly over the threads.
__device__ __forceinline__ void clear_spans(const GridGeom& g, const MoveResult& mr,
                                            const LayerTable& lt, int policy, size_t tid,
                                            size_t nthreads) {
  const uint32_t rows_local = static_cast<uint32_t>(g.row_end - g.row_begin);
  const size_t cells = static_cast<size_t>(rows_local) * g.cols;
  const int n_layers = (policy == 1) ? 3 : lt.count;
  if (mr.clear_all) {
    for (int li = 0; li < n_layers; ++li) {
      float* __restrict__ d = lt.ptr[(policy == 1) ? lt.basic[li] : li];
      for (size_t c = tid; c < cells; c += nthreads) d[c] = nan_f32();
    }
    return;
  }
  uint32_t span_size[4] = {0, 0, 0, 0};
  uint32_t per_layer = 0;
  for (int sidx = 0; sidx < mr.n_spans; ++sidx) {
    const ClearSpan sp = mr.spans[sidx];
    span_size[sidx] = static_cast<uint32_t>(sp.n) * (sp.axis == 0 ? static_cast<uint32_t>(g.cols) : rows_local);
    per_layer += span_size[sidx];
  }
  if (per_layer == 0) return;
  const size_t total = static_cast<size_t>(per_layer) * n_layers;
  for (size_t w = tid; w < total; w += nthreads) {
    const uint32_t li = static_cast<uint32_t>(w / per_layer);
    uint32_t t = static_cast<uint32_t>(w - static_cast<size_t>(li) * per_layer);
    int sidx = 0;
    while (t >= span_size[sidx]) {
      t -= span_size[sidx];
      ++sidx;
    }
    const ClearSpan sp = mr.spans[sidx];
    float* __restrict__ d = lt.ptr[(policy == 1) ? lt.basic[li] : li];
    if (sp.axis == 0) {  // buffer rows [k, k+n) of every column (LOCAL maps are unsharded)
      const uint32_t c = t / static_cast<uint32_t>(sp.n);
      const uint32_t r = static_cast<uint32_t>(sp.k) + (t - c * static_cast<uint32_t>(sp.n));
      d[static_cast<size_t>(c) * rows_local + r] = nan_f32();
    } else {             // buffer columns [k, k+n): contiguous in column-major storage
      d[static_cast<size_t>(sp.k) * rows_local + t] = nan_f32();
    }
  }
}

__global__ void __launch_bounds__(kBlock)
commit_move_clear_kernel(const __grid_constant__ CommitParams p,
                         const DeviceState* __restrict__ st_in, DeviceState* __restrict__ st_out,
                         uint32_t* __restrict__ counters,
                         const __grid_constant__ LayerTable lt) {
  // The three jobs of this kernel are independent, so the grid is split into role groups
  // that run them side by side (each job is a short chain of dependent memory round trips;
  // doing them one after another in every thread would add the chains up):
  //   group A  bucket segment allocation (tile path)
  //   group B  reset of the last observing scan's obstacle cells
  //   group C  state commit + circular-buffer move (vacated rows / columns -> NaN)
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();  // K1's counters / bucket histogram are complete and visible from here on
  const uint32_t nA = p.tile_path ? gridDim.x / 2 : 0;
  const uint32_t nB = p.tile_path ? 0 : gridDim.x / 2;  // tile path: the scatter grid does job B
  const uint32_t s_inside = counters[CNT_INSIDE];

  if (blockIdx.x < nA) {
    // tile path, L1 segment allocation: every non-empty bucket gets a contiguous run of
    // record slots sized by its point count (an upper bound on its records) and joins the
    // work list.  Where a bucket lands is irrelevant to the result, so no global prefix
    // scan is needed — two atomics per WARP of 32 buckets, on scan scratch.
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t nthreads = static_cast<size_t>(nA) * blockDim.x;
    for (size_t b = tid; b - lane < p.tb.n_buckets; b += nthreads) {
      const uint32_t cnt = b < p.tb.n_buckets ? p.tb.bucket_count[b] : 0u;
      uint32_t incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      const uint32_t nz = __ballot_sync(0xffffffffu, cnt != 0);
      uint32_t slot0 = 0, list0 = 0;
      if (lane == 0 && total) {
        slot0 = atomicAdd(&counters[CNT_REC_SLOTS], total);
        list0 = atomicAdd(&counters[CNT_BUCKETS], __popc(nz));
      }
      slot0 = __shfl_sync(0xffffffffu, slot0, 0);
      list0 = __shfl_sync(0xffffffffu, list0, 0);
      if (cnt) {
        const uint32_t off = slot0 + incl - cnt;
        p.tb.bucket_offset[b] = off;
        p.tb.bucket_list[list0 + __popc(nz & ((1u << lane) - 1u))] =
            make_uint4(static_cast<uint32_t>(b), off, cnt, 0u);
      }
    }
    return;
  }

  if (blockIdx.x < nA + nB) {
    // updateObstacle's map_.clear(obstacle) (elevation_mapping.cpp:146) restricted to the
    // cells that can hold a value: those the last observing scan touched.  Runs only when
    // this scan has observations (update() returns early otherwise, :116-117).
    if (s_inside > 0 && p.obstacle) {
      const size_t tid = static_cast<size_t>(blockIdx.x - nA) * blockDim.x + threadIdx.x;
      reset_obstacle(p.obstacle, p.touched_keys, st_in->touched_count, p.invalid_key,
                     (st_in->flags & SF_OBSTACLE_DIRTY) != 0, p.obstacle_cells, tid,
                     static_cast<size_t>(nB) * blockDim.x);
    }
    return;
  }

  __shared__ GridGeom g_new;
  __shared__ MoveResult mr;
  if (threadIdx.x == 0) {
    const GridGeom g_old = st_in->geom;
    const uint32_t kept = counters[CNT_KEPT];
    mr.moved = 0;
    mr.clear_all = 0;
    mr.n_spans = 0;
    g_new = g_old;
    // map_.move() runs only when preprocessScan left >= 1 point (fastdem.cpp:137-138)
    if (p.local_mode && kept > 0) g_new = geom_move(g_old, p.robot_x, p.robot_y, mr);
    if (blockIdx.x == nA + nB) {
      st_out->geom = g_new;
      // sticky facts (StateFlag): layers the reference has created lazily by now; the
      // caller-edited obstacle layer is clean again once an observing scan has cleared it
      uint32_t flags = st_in->flags;
      if (s_inside > 0) flags = (flags | p.flags_if_cells) & ~static_cast<uint32_t>(SF_OBSTACLE_DIRTY);
      if (p.raycast && kept > 0 && geom_is_inside(g_new, p.rc_origin_x, p.rc_origin_y)) flags |= SF_RAYCAST;
      st_out->flags = flags;
      if (p.defer) {
        // batched integration: the map writes (and the touched-count hand-over, which the
        // previous scan's estimator may still be producing) belong to back_prologue_kernel
        p.move_out->mr = mr;
        p.move_out->g_new = g_new;
      } else {
        // The touched list is replaced only when this scan produced observations.  Global
        // sort path: K3 writes one slot per sorted element (invalid_key except at segment
        // tails).  Tile path: K3t appends compactly and counts up from 0.
        if (s_inside > 0) st_out->touched_count = p.tile_path ? 0u : s_inside;
        else st_out->touched_count = st_in->touched_count;
      }
    }
  }
  __syncthreads();
  if (!p.defer && (mr.clear_all || mr.n_spans > 0)) {
    const uint32_t nC = gridDim.x - nA - nB;
    const size_t tid = static_cast<size_t>(blockIdx.x - nA - nB) * blockDim.x + threadIdx.x;
    clear_spans(g_new, mr, lt, p.clear_policy, tid, static_cast<size_t>(nC) * blockDim.x);
  }
}

// Batched integration, first kernel of a scan's BACK half (runs after the previous scan's
// estimator; the front half — K1, commit, scatter — of this scan ran beside it):
//   * touched-count hand-over: this scan's list starts empty if it has observations
//   * updateObstacle's map_.clear(obstacle) restricted to the previous touched cells
//   * GridMap::move()'s clearing of the vacated rows / columns recorded by the commit
__global__ void __launch_bounds__(kBlock)
back_prologue_kernel(const __grid_constant__ BackParams p, const __grid_constant__ LayerTable lt) {
  pdl_launch_dependents();  // K3t of this scan may set up (shared-memory init) while this grid runs
  pdl_wait();               // the previous scan's K3t is complete and flushed from here on
  const uint32_t n_inside = p.counters[CNT_INSIDE];
  const uint32_t prev = p.st_cur->touched_count;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.st_out->touched_count = n_inside > 0 ? 0u : prev;
  const uint32_t nB = gridDim.x / 2;
  if (blockIdx.x < nB) {
    if (n_inside > 0 && p.obstacle) {
      const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
      reset_obstacle(p.obstacle, p.touched_keys, prev, p.invalid_key,
                     (p.st_cur->flags & SF_OBSTACLE_DIRTY) != 0, p.obstacle_cells, tid,
                     static_cast<size_t>(nB) * blockDim.x);
    }
    return;
  }
  __shared__ MoveRecord rec;
  if (threadIdx.x == 0) rec = *p.move;
  __syncthreads();
  if (rec.mr.clear_all || rec.mr.n_spans > 0) {
    const size_t tid = static_cast<size_t>(blockIdx.x - nB) * blockDim.x + threadIdx.x;
    clear_spans(rec.g_new, rec.mr, lt, p.clear_policy, tid, static_cast<size_t>(gridDim.x - nB) * blockDim.x);
  }
}

// End of a scan: hand the scan statistics + committed state to the host (pinned, mapped
// memory — no memcpy nodes), make the committed state current, and re-arm the counters for
// the next scan (no memset node).  One warp.
__global__ void publish_kernel(uint32_t* __restrict__ counters, DeviceState* __restrict__ st_cur,
                               const DeviceState* __restrict__ st_next,
                               uint32_t* __restrict__ host_out) {
  constexpr int kStateWords = sizeof(DeviceState) / 4;
  static_assert(sizeof(DeviceState) % 4 == 0 && kStateWords <= 32 && CNT_COUNT <= 32, "one warp");
  const int t = threadIdx.x;
  uint32_t c = 0, w = 0;
  if (t < CNT_COUNT) c = counters[t];
  if (t < kStateWords) w = reinterpret_cast<const uint32_t*>(st_next)[t];
  if (t < CNT_COUNT) {
    host_out[t] = c;
    counters[t] = 0;
  }
  if (t < kStateWords) {
    host_out[CNT_COUNT + t] = w;
    reinterpret_cast<uint32_t*>(st_cur)[t] = w;
  }
}

// GridMap::move() on its own (fdem_map_move)
__global__ void __launch_bounds__(kBlock)
move_only_kernel(const DeviceState* __restrict__ st_in, DeviceState* __restrict__ st_out, double x,
                 double y, int clear_policy, const __grid_constant__ LayerTable lt,
                 uint32_t* moved_flag) {
  __shared__ GridGeom g_new;
  __shared__ MoveResult mr;
  if (threadIdx.x == 0) {
    g_new = geom_move(st_in->geom, x, y, mr);
    if (blockIdx.x == 0) {
      st_out->geom = g_new;
      st_out->touched_count = st_in->touched_count;
      st_out->flags = st_in->flags;
      *moved_flag = mr.moved;
    }
  }
  __syncthreads();
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  if (mr.clear_all || mr.n_spans > 0) clear_spans(g_new, mr, lt, clear_policy, tid, nthreads);
}

// ───────────────────────────── K3: segmented reduce + estimator ──────────────
// Global-sort path.  One warp owns every segment (cell) whose first sorted element lies
// in its 32-element window; it follows a segment past the window if needed, 32 elements
// per step, so a cell is always reduced and written by exactly one lane of exactly one
// warp.  Reduction = segmented inclusive scan over warp shuffles.
__global__ void __launch_bounds__(kBlock)
segreduce_estimate_kernel(const __grid_constant__ EstimateParams p,
                          const uint32_t* __restrict__ counters_ro,
                          uint32_t* __restrict__ counters) {
  const uint32_t n_valid = min(counters_ro[CNT_INSIDE], p.n_sorted);
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t chunk_begin64 = static_cast<uint64_t>(warp_global) * 32u;
  if (chunk_begin64 >= n_valid) return;
  const uint32_t chunk_begin = static_cast<uint32_t>(chunk_begin64);
  const uint32_t chunk_end = min(chunk_begin + 32u, n_valid);
  const uint32_t INV = p.invalid_key;
  const bool has_i = p.intensity != nullptr;

  bool open = false;  // an owned segment is open at the start of the window (warp-uniform)
  CellObs carry = obs_identity();
  uint32_t cells_done = 0;

  for (uint32_t base = chunk_begin;; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < n_valid;
    // keys and point indices are independent loads: issue both before anything depends on them
    const uint32_t k = valid ? __ldg(&p.sorted_keys[i]) : INV;
    const uint32_t idx = valid ? __ldg(&p.sorted_vals[i]) : 0u;
    uint32_t kprev = __shfl_up_sync(0xffffffffu, k, 1);
    if (lane == 0) kprev = (i > 0 && valid) ? __ldg(&p.sorted_keys[i - 1]) : INV;
    uint32_t knext = __shfl_down_sync(0xffffffffu, k, 1);
    if (lane == 31) knext = (i + 1 < n_valid) ? __ldg(&p.sorted_keys[i + 1]) : INV;
    const bool head = valid && (i == 0 || k != kprev);
    const bool tail = valid && (i + 1 >= n_valid || k != knext);

    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t m = heads & (0xffffffffu >> (31 - lane));  // heads at lanes <= mine
    const bool head_in_win = m != 0;
    const int s = head_in_win ? (31 - __clz(m)) : 0;          // lane where my segment starts
    const bool owned = valid && (head_in_win ? (base + s < chunk_end) : open);

    CellObs v = obs_identity();
    if (owned) {
      const float4 q = __ldg(&p.pm[idx]);
      const float in = has_i ? __ldg(&p.intensity[idx]) : 0.0f;
      v = obs_from_point(q.z, q.w, in, has_i, idx);
    }

    // segmented inclusive scan within the window (Hillis-Steele over shuffles)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const CellObs o = obs_shfl_up(v, d);
      if (lane - d >= s) v = obs_combine(o, v);
    }
    // my segment started in an earlier window: fold the carried prefix in
    if (!head_in_win && open) v = obs_combine(carry, v);

    // touched-cell list for the next scan's obstacle reset: the key sits at the segment's
    // TAIL position (where min_z is known), invalid_key everywhere else
    if (owned) p.touched_keys[i] = tail ? k : INV;

    if (owned && tail) {
      ++cells_done;
      apply_observation(p, k, v);
      if (p.touched_minz) p.touched_minz[i] = v.mz;
    }

    // carry the open segment (if any) into the next window
    const uint32_t remaining = n_valid - base;  // >= 1
    const int last_lane = remaining >= 32 ? 31 : static_cast<int>(remaining) - 1;
    const bool last_tail = __shfl_sync(0xffffffffu, tail, last_lane);
    const bool last_owned = __shfl_sync(0xffffffffu, owned, last_lane);
    carry = obs_shfl(v, last_lane);
    open = last_owned && !last_tail;
    if (remaining <= 32) break;                     // stream exhausted
    if (base + 32 >= chunk_end && !open) break;     // nothing of mine continues
  }

  // scan statistic (n_cells), one atomic per warp
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) cells_done += __shfl_down_sync(0xffffffffu, cells_done, d);
  if (lane == 0 && cells_done) atomicAdd(&counters[CNT_CELLS], cells_done);
}

inline int grid_for(size_t n, int block, int max_blocks = 148 * 8) {
  size_t b = (n + block - 1) / block;
  if (b < 1) b = 1;
  if (b > static_cast<size_t>(max_blocks)) b = max_blocks;
  return static_cast<int>(b);
}

}  // namespace

// ───────────────────────────── launchers ─────────────────────────────────────

void launch_fill(float* dst, size_t n, float v, cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  fill_kernel<<<grid_for(n, kBlock), kBlock, 0, s>>>(dst, n, v);
  ++lc.mine;
}
void launch_fill_u32(uint32_t* dst, size_t n, uint32_t v, cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  fill_u32_kernel<<<grid_for(n, kBlock), kBlock, 0, s>>>(dst, n, v);
  ++lc.mine;
}
void launch_any_not_nan(const float* src, size_t n, uint32_t* flag, cudaStream_t s,
                        LaunchCounter& lc) {
  if (n == 0) return;
  any_not_nan_kernel<<<grid_for(n, kBlock), kBlock, 0, s>>>(src, n, flag);
  ++lc.mine;
}
void launch_clear_cell(const LayerTable& lt, int64_t lin, cudaStream_t s, LaunchCounter& lc) {
  clear_cell_kernel<<<1, 64, 0, s>>>(lt, lin);
  ++lc.mine;
}
void launch_preprocess_bin(const PreprocessParams& p, const DeviceState* st_in, uint32_t* counters,
                           float4* pm, uint32_t* keys, uint32_t* vals, cudaStream_t s,
                           LaunchCounter& lc) {
  if (p.n == 0) return;
  const int grid = static_cast<int>((p.n + kBlock - 1) / kBlock);
  preprocess_bin_kernel<<<grid, kBlock, 0, s>>>(p, st_in, counters, pm, keys, vals);
  ++lc.mine;
}
void launch_commit(const CommitParams& p, const DeviceState* st_in, DeviceState* st_out,
                   uint32_t* counters, const LayerTable& lt, cudaStream_t s, LaunchCounter& lc) {
  commit_move_clear_kernel<<<148, kBlock, 0, s>>>(p, st_in, st_out, counters, lt);
  ++lc.mine;
}
void launch_publish(uint32_t* counters, DeviceState* st_cur, const DeviceState* st_next,
                    uint32_t* host_out, cudaStream_t s, LaunchCounter& lc) {
  publish_kernel<<<1, 32, 0, s>>>(counters, st_cur, st_next, host_out);
  ++lc.mine;
}
void launch_segreduce_estimate(const EstimateParams& p, const uint32_t* counters_ro,
                               uint32_t* counters, cudaStream_t s, LaunchCounter& lc) {
  if (p.n_sorted == 0) return;
  const size_t warps = (static_cast<size_t>(p.n_sorted) + 31) / 32;
  const int grid = static_cast<int>((warps * 32 + kBlock - 1) / kBlock);
  segreduce_estimate_kernel<<<grid, kBlock, 0, s>>>(p, counters_ro, counters);
  ++lc.mine;
}
void launch_move_only(const DeviceState* st_in, DeviceState* st_out, double x, double y,
                      int clear_policy, const LayerTable& lt, uint32_t* moved_flag, cudaStream_t s,
                      LaunchCounter& lc) {
  move_only_kernel<<<148, kBlock, 0, s>>>(st_in, st_out, x, y, clear_policy, lt, moved_flag);
  ++lc.mine;
}

}  // namespace fdem

// ── launch descriptors for the per-scan CUDA graph (capi.cu packs the arguments) ──
namespace fdem {
KernelDesc desc_preprocess_bin(uint32_t n) {
  return KernelDesc{reinterpret_cast<const void*>(&preprocess_bin_kernel),
                    dim3((n + kBlock - 1) / kBlock), dim3(kBlock), 0};
}
KernelDesc desc_commit() {
  return KernelDesc{reinterpret_cast<const void*>(&commit_move_clear_kernel), dim3(148),
                    dim3(kBlock), 0};
}
KernelDesc desc_back_prologue() {
  // a small grid: the kernel sits on the batch's serial back chain and its work is tens of
  // thousands of independent stores at most (FDEM_BP_CTAS overrides, for tuning)
  static const int ctas = [] {
    const char* e = std::getenv("FDEM_BP_CTAS");
    const int v = e ? std::atoi(e) : 148;
    return v >= 2 && v <= 1184 ? v : 148;
  }();
  return KernelDesc{reinterpret_cast<const void*>(&back_prologue_kernel), dim3(ctas), dim3(kBlock), 0};
}
KernelDesc desc_publish() {
  return KernelDesc{reinterpret_cast<const void*>(&publish_kernel), dim3(1), dim3(32), 0};
}
}  // namespace fdem
