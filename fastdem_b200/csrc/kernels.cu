// kernels.cu — the integrate() hot path as sm_100a CUDA kernels.
//
//   K1 preprocess_bin_kernel      sensor covariance + SE(3) x2 + range/height crop +
//                                 sigma_z^2 = (R S R^T)(2,2) + FP64 (x,y) -> cell key
//   K2 commit_move_clear_kernel   LOCAL-mode circular-buffer move (vacated stripes -> NaN),
//                                 reset of last scan's obstacle cells, scan-state commit
//   (sort by cell: sort.cu)
//   K3 segreduce_estimate_kernel  warp-segmented min/max reduce over the sorted stream +
//                                 one Kalman / P2 state step per touched cell, every layer
//                                 written once, no global atomics on estimator state
//
// Compiled with -fmad=false: every float/double expression below is evaluated exactly as
// written (no FMA contraction), in the operation order the CPU oracle fixes, so cell
// indices are bit-identical and heights/variances agree to the last bit in practice.
// Reference file:line citations are relative to /root/reference/.
#include <float.h>
#include <math.h>

#include "device_types.h"

namespace fdem {

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ float nanf_() { return __int_as_float(0x7fc00000); }

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
  if (l < lt.count) lt.ptr[l][lin] = nanf_();
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

  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool kept = false, inside = false;
  if (i < p.n) {
    float4 q = __ldg(&p.xyzw[i]);
    float var_z = 0.0f;
    if (p.input_frame == INPUT_SENSOR_FRAME) {
      // preprocessScan (fastdem/src/fastdem.cpp:164-190)
      float S[9];
      if (p.cov9) {
#pragma unroll
        for (int k = 0; k < 9; ++k) S[k] = __ldg(&p.cov9[static_cast<size_t>(i) * 9 + k]);
      } else {
        sensor_covariance(p, q, S);
      }
      q = transform_point(p.T1, q);                       // sensor -> base
      const float d2 = sqnorm3(q.x, q.y, q.z);            // cropRange, base frame
      kept = (d2 >= p.range_min_sq && d2 <= p.range_max_sq) &&
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

    uint32_t key = p.invalid_key;
    if (kept) {
      int32_t row, col;
      if (geom_get_index(sg, static_cast<double>(q.x), static_cast<double>(q.y), row, col)) {
        const int64_t lin = geom_linear(sg, row, col);
        if (lin >= 0) {
          key = static_cast<uint32_t>(lin);
          inside = true;
        }
      }
      pm[i] = make_float4(q.x, q.y, q.z, var_z);
    } else {
      pm[i] = make_float4(nanf_(), nanf_(), nanf_(), 0.0f);  // dropped by the crop filters
    }
    keys[i] = key;
    vals[i] = i;
  }

  // block-level counts -> one atomic per block per counter (scan statistics, not map state)
  const uint32_t kept_w = __popc(__ballot_sync(0xffffffffu, kept));
  const uint32_t inside_w = __popc(__ballot_sync(0xffffffffu, inside));
  if ((threadIdx.x & 31) == 0) {
    if (kept_w) atomicAdd(&s_kept, kept_w);
    if (inside_w) atomicAdd(&s_inside, inside_w);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_kept) atomicAdd(&counters[CNT_KEPT], s_kept);
    if (s_inside) atomicAdd(&counters[CNT_INSIDE], s_inside);
  }
}

// ───────────────────────────── K2: commit / move / clear ─────────────────────

__device__ __forceinline__ void clear_spans(const GridGeom& g, const MoveResult& mr,
                                            const LayerTable& lt, int policy, size_t tid,
                                            size_t nthreads) {
  const int rows_local = g.row_end - g.row_begin;
  const size_t cells = static_cast<size_t>(rows_local) * g.cols;
  const int n_layers = (policy == 1) ? 3 : lt.count;
  for (int li = 0; li < n_layers; ++li) {
    float* __restrict__ d = lt.ptr[(policy == 1) ? lt.basic[li] : li];
    if (mr.clear_all) {
      for (size_t c = tid; c < cells; c += nthreads) d[c] = nanf_();
      continue;
    }
    for (int sidx = 0; sidx < mr.n_spans; ++sidx) {
      const ClearSpan sp = mr.spans[sidx];
      if (sp.axis == 0) {  // buffer rows [k, k+n) of every column (LOCAL maps are unsharded)
        const size_t total = static_cast<size_t>(sp.n) * g.cols;
        for (size_t t = tid; t < total; t += nthreads) {
          const size_t c = t / sp.n;
          const size_t r = sp.k + (t - c * sp.n);
          d[c * rows_local + r] = nanf_();
        }
      } else {  // buffer columns [k, k+n): contiguous in column-major storage
        const size_t total = static_cast<size_t>(sp.n) * rows_local;
        float* __restrict__ base = d + static_cast<size_t>(sp.k) * rows_local;
        for (size_t t = tid; t < total; t += nthreads) base[t] = nanf_();
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock)
commit_move_clear_kernel(const __grid_constant__ CommitParams p,
                         const DeviceState* __restrict__ st_in, DeviceState* __restrict__ st_out,
                         const uint32_t* __restrict__ counters,
                         const __grid_constant__ LayerTable lt) {
  __shared__ GridGeom g_new;
  __shared__ MoveResult mr;
  __shared__ uint32_t s_inside, s_prev_touched;
  if (threadIdx.x == 0) {
    const GridGeom g_old = st_in->geom;
    const uint32_t kept = counters[CNT_KEPT];
    s_inside = counters[CNT_INSIDE];
    s_prev_touched = st_in->touched_count;
    mr.moved = 0;
    mr.clear_all = 0;
    mr.n_spans = 0;
    g_new = g_old;
    // map_.move() runs only when preprocessScan left >= 1 point (fastdem.cpp:137-138)
    if (p.local_mode && kept > 0) g_new = geom_move(g_old, p.robot_x, p.robot_y, mr);
    if (blockIdx.x == 0) {
      st_out->geom = g_new;
      // the touched list is replaced by K3 only when this scan produced observations
      st_out->touched_count = s_inside > 0 ? s_inside : s_prev_touched;
    }
  }
  __syncthreads();
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;

  if (mr.clear_all || mr.n_spans > 0) clear_spans(g_new, mr, lt, p.clear_policy, tid, nthreads);

  // updateObstacle's map_.clear(obstacle) (elevation_mapping.cpp:146) restricted to the
  // cells that can hold a value: those the last observing scan touched.  Runs only when
  // this scan has observations (update() returns early otherwise, :116-117).
  if (s_inside > 0 && p.obstacle) {
    for (size_t j = tid; j < s_prev_touched; j += nthreads) {
      const uint32_t k = p.touched_keys[j];
      if (k != p.invalid_key) p.obstacle[k] = nanf_();
    }
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
      *moved_flag = mr.moved;
    }
  }
  __syncthreads();
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  if (mr.clear_all || mr.n_spans > 0) clear_spans(g_new, mr, lt, clear_policy, tid, nthreads);
}

// ───────────────────────────── K3: segmented reduce + estimator ──────────────

// per-scan observation of one cell (ElevationMapping::CellObservation,
// mapping/elevation_mapping.hpp:26-34) as carried through the segmented scan
struct Obs {
  float mz;  // min_z        (init FLT_MAX)
  float mv;  // min_z_var    (variance of the FIRST point attaining min_z)
  float xz;  // max_z        (init -FLT_MAX)
  float it;  // max intensity over non-NaN values (init -inf)
};

// left-biased combine: `a` precedes `b` in point-index order (the sort is stable), so a
// strict `<` keeps the lowest-index point on equal min_z — rasterize()'s `z < cell.min_z`
// (elevation_mapping.cpp:65-68)
__device__ __forceinline__ Obs combine(const Obs& a, const Obs& b) {
  Obs r;
  const bool take_b = b.mz < a.mz;
  r.mz = take_b ? b.mz : a.mz;
  r.mv = take_b ? b.mv : a.mv;
  r.xz = (b.xz > a.xz) ? b.xz : a.xz;
  r.it = (b.it > a.it) ? b.it : a.it;
  return r;
}

__device__ __forceinline__ Obs shfl_up_obs(const Obs& v, int d) {
  Obs r;
  r.mz = __shfl_up_sync(0xffffffffu, v.mz, d);
  r.mv = __shfl_up_sync(0xffffffffu, v.mv, d);
  r.xz = __shfl_up_sync(0xffffffffu, v.xz, d);
  r.it = __shfl_up_sync(0xffffffffu, v.it, d);
  return r;
}
__device__ __forceinline__ Obs shfl_obs(const Obs& v, int src) {
  Obs r;
  r.mz = __shfl_sync(0xffffffffu, v.mz, src);
  r.mv = __shfl_sync(0xffffffffu, v.mv, src);
  r.xz = __shfl_sync(0xffffffffu, v.xz, src);
  r.it = __shfl_sync(0xffffffffu, v.it, src);
  return r;
}

// Kalman::update + computeBounds on one cell (mapping/kalman_estimation.hpp:98-153)
__device__ __forceinline__ void kalman_cell(const EstimateParams& p, uint32_t c, float z,
                                            float meas_var) {
  const EstLayers& L = p.L;
  float x = L.elevation[c];
  float P = L.kalman_p[c];
  float count = L.n_points[c];
  float mean = L.sample_mean[c];
  float svar = L.variance[c];
  float m2 = L.sample_m2[c];

  const float R = (meas_var > 0.0f) ? meas_var : p.kalman_max_variance;
  if (isnan(x)) {
    x = z;
    P = R;
    count = 1.0f;
  } else {
    P += p.kalman_process_noise;
    const float K = P / (P + R);
    x = x + K * (z - x);
    P = (1.0f - K) * P;
    P = fminf(fmaxf(P, p.kalman_min_variance), p.kalman_max_variance);
    count += 1.0f;
  }
  if (isnan(mean)) {
    mean = z;
    svar = 0.0f;
    m2 = 0.0f;
  } else {
    const float delta = z - mean;
    const float new_mean = mean + (delta / count);
    const float delta2 = z - new_mean;
    m2 += delta * delta2;
    svar = (count > 1.0f) ? m2 / (count - 1.0f) : 0.0f;
    mean = new_mean;
  }
  const float sigma = sqrtf(fmaxf(0.0f, svar));
  L.elevation[c] = x;
  L.kalman_p[c] = P;
  L.n_points[c] = count;
  L.sample_mean[c] = mean;
  L.variance[c] = svar;
  L.sample_m2[c] = m2;
  L.upper_bound[c] = x + 2.0f * sigma;
  L.lower_bound[c] = x - 2.0f * sigma;
}

__device__ __forceinline__ float p2_parabolic(const float* q, const float* n, int i, int sign) {
  const float d_right = n[i + 1] - n[i];
  const float d_left = n[i] - n[i - 1];
  const float d_span = n[i + 1] - n[i - 1];
  if (d_right == 0.0f || d_left == 0.0f || d_span == 0.0f) return q[i];
  const float s = static_cast<float>(sign);
  const float t1 = (d_left + s) * (q[i + 1] - q[i]) / d_right;
  const float t2 = (d_right - s) * (q[i] - q[i - 1]) / d_left;
  return q[i] + s * (t1 + t2) / d_span;
}
__device__ __forceinline__ float p2_linear(const float* q, const float* n, int i, int sign) {
  const int j = i + sign;
  const float dn = n[j] - n[i];
  if (dn == 0.0f) return q[i];
  return q[i] + static_cast<float>(sign) * (q[j] - q[i]) / dn;
}

// P2Quantile::update + updateP2 + computeBounds on one cell
// (mapping/quantile_estimation.hpp:141-258)
__device__ __forceinline__ void p2_cell(const EstimateParams& p, uint32_t c, float x) {
  const EstLayers& L = p.L;
  float q[5], n[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    q[k] = L.p2_q[k][c];
    n[k] = L.p2_n[k][c];
  }
  float count = L.n_points[c];
  if (isnan(count) || count < 0.0f) count = 0.0f;
  if (count < 5.0f) {
    // phase 1: collect the first five samples
    const int slot = static_cast<int>(count);
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (k == slot) q[k] = x;
    count += 1.0f;
    if (count >= 5.0f) {
      // std::sort(q, q+5): insertion sort, as libstdc++ does below 16 elements
#pragma unroll
      for (int i = 1; i < 5; ++i) {
        const float v = q[i];
        int j = i - 1;
        while (j >= 0 && v < q[j]) {
          q[j + 1] = q[j];
          --j;
        }
        q[j + 1] = v;
      }
#pragma unroll
      for (int i = 0; i < 5; ++i) n[i] = static_cast<float>(i);
    }
  } else {
    int k;
    if (x < q[0]) {
      q[0] = x;
      k = 0;
    } else if (x < q[1]) {
      k = 0;
    } else if (x < q[2]) {
      k = 1;
    } else if (x < q[3]) {
      k = 2;
    } else if (x <= q[4]) {
      k = 3;
    } else {
      q[4] = x;
      k = 3;
    }
#pragma unroll
    for (int i = 1; i < 5; ++i)
      if (i > k) n[i] += 1.0f;
    float n_prime[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) n_prime[i] = p.p2_dn[i] * count;  // pre-increment count (:216-219)
    count += 1.0f;
    if (p.p2_max_sample_count > 0.0f && count > p.p2_max_sample_count) {
      const float scale = p.p2_max_sample_count / count;
#pragma unroll
      for (int i = 0; i < 5; ++i) n[i] *= scale;
      count = p.p2_max_sample_count;
    }
#pragma unroll
    for (int i = 1; i < 4; ++i) {
      const float d = n_prime[i] - n[i];
      if ((d >= 1.0f && n[i + 1] - n[i] > 1.0f) || (d <= -1.0f && n[i - 1] - n[i] < -1.0f)) {
        const int sign = (d >= 0.0f) ? 1 : -1;
        const float q_new = p2_parabolic(q, n, i, sign);
        q[i] = (q[i - 1] < q_new && q_new < q[i + 1]) ? q_new : p2_linear(q, n, i, sign);
        n[i] += static_cast<float>(sign);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    L.p2_q[k][c] = q[k];
    L.p2_n[k][c] = n[k];
  }
  L.n_points[c] = count;
  // update() writes (count>=5 ? q[m] : x) but computeBounds() immediately overwrites
  // elevation with q[m] (:161-162, :172) — only the latter survives estimate().
  float qm = q[0];
#pragma unroll
  for (int k = 1; k < 5; ++k)
    if (k == p.p2_marker) qm = q[k];
  L.elevation[c] = qm;
  const float sigma = (q[3] - q[1]) / 2.0f;
  L.variance[c] = sigma * sigma;
  L.lower_bound[c] = q[0];
  L.upper_bound[c] = q[4];
}

constexpr int kK3Windows = 4;                  // 32-element windows per warp chunk
constexpr int kK3Chunk = 32 * kK3Windows;      // sorted elements whose segment HEADS a warp owns

// One warp owns every segment (cell) whose first sorted element lies in its chunk; it
// follows a segment past the chunk end if needed, 32 elements per step, so a cell is
// always reduced and written by exactly one lane of exactly one warp.
__global__ void __launch_bounds__(kBlock)
segreduce_estimate_kernel(const __grid_constant__ EstimateParams p,
                          const uint32_t* __restrict__ counters_ro,
                          uint32_t* __restrict__ counters) {
  const uint32_t n_valid = min(counters_ro[CNT_INSIDE], p.n_sorted);
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t chunk_begin64 = static_cast<uint64_t>(warp_global) * kK3Chunk;
  if (chunk_begin64 >= n_valid) return;
  const uint32_t chunk_begin = static_cast<uint32_t>(chunk_begin64);
  const uint32_t chunk_end = min(chunk_begin + static_cast<uint32_t>(kK3Chunk), n_valid);
  const uint32_t INV = p.invalid_key;
  const bool has_i = p.intensity != nullptr;
  const bool has_c = p.rgb != nullptr;

  bool open = false;  // an owned segment is open at the start of the window (warp-uniform)
  Obs carry;
  carry.mz = FLT_MAX; carry.mv = 0.0f; carry.xz = -FLT_MAX; carry.it = -INFINITY;
  bool carry_first_nan = false;
  uint32_t cells_done = 0;

  for (uint32_t base = chunk_begin;; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < n_valid;
    const uint32_t k = valid ? __ldg(&p.sorted_keys[i]) : INV;
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

    Obs v;
    v.mz = FLT_MAX; v.mv = 0.0f; v.xz = -FLT_MAX; v.it = -INFINITY;
    bool my_nan = false;
    uint32_t idx = 0;
    if (owned) {
      idx = __ldg(&p.sorted_vals[i]);
      const float4 q = __ldg(&p.pm[idx]);
      // rasterize() folds from {FLT_MAX, 0, lowest}: a point only enters through a strict
      // compare, so NaN / out-of-range z leave the initial values in place
      if (q.z < FLT_MAX) { v.mz = q.z; v.mv = q.w; }
      if (q.z > -FLT_MAX) v.xz = q.z;
      if (has_i) {
        const float in = __ldg(&p.intensity[idx]);
        my_nan = isnan(in);
        if (!my_nan) v.it = in;
      }
    }

    // segmented inclusive scan within the window (Hillis-Steele over shuffles)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const Obs o = shfl_up_obs(v, d);
      if (lane - d >= s) v = combine(o, v);
    }
    bool first_nan = __shfl_sync(0xffffffffu, my_nan, s);
    if (!head_in_win) {
      // my segment started in an earlier window: fold the carried prefix in from the left
      if (open) v = combine(carry, v);
      first_nan = carry_first_nan;
    }

    // touched-cell list for the next scan's obstacle reset: the key sits at the segment's
    // TAIL position (where min_z is known), invalid_key everywhere else
    if (owned) p.touched_keys[i] = tail ? k : INV;

    if (owned && tail) {
      ++cells_done;
      const uint32_t c = k;
      // estimate(): estimator.update(idx, min_z, min_z_var) + computeBounds (elevation_mapping.cpp:94-108)
      if (p.estimation_type == 1) p2_cell(p, c, v.mz);
      else kalman_cell(p, c, v.mz, v.mv);
      // updateMinMax (elevation_mapping.cpp:127-142)
      const float smin = p.L.elevation_min[c];
      if (isnan(smin) || v.mz < smin) p.L.elevation_min[c] = v.mz;
      const float smax = p.L.elevation_max[c];
      if (isnan(smax) || v.xz > smax) p.L.elevation_max[c] = v.xz;
      // updateObstacle (elevation_mapping.cpp:144-152)
      p.L.obstacle[c] = (v.xz > v.mz) ? v.xz : nanf_();
      // updateIntensity (elevation_mapping.cpp:154-166): the per-scan max is NaN only when the
      // first point of the cell carries NaN (rasterize :72-78 takes the first value blindly)
      if (has_i) {
        const float mi = first_nan ? nanf_() : v.it;
        const float stored = p.L.intensity[c];
        if (isnan(stored) || mi > stored) p.L.intensity[c] = mi;
      }
      // updateColor (elevation_mapping.cpp:168-175): last point of the cell wins; the tail
      // lane IS the last point (stable sort).  0x00RRGGBB reinterpreted as float.
      if (has_c) {
        const uint8_t* rgb = p.rgb + static_cast<size_t>(idx) * 3;
        const uint32_t bits = (static_cast<uint32_t>(rgb[0]) << 16) |
                              (static_cast<uint32_t>(rgb[1]) << 8) | rgb[2];
        reinterpret_cast<uint32_t*>(p.L.color)[c] = bits;
      }
      if (p.touched_minz) p.touched_minz[i] = v.mz;
    }

    // carry the open segment (if any) into the next window
    const uint32_t remaining = n_valid - base;  // >= 1
    const int last_lane = remaining >= 32 ? 31 : static_cast<int>(remaining) - 1;
    const bool last_tail = __shfl_sync(0xffffffffu, tail, last_lane);
    const bool last_owned = __shfl_sync(0xffffffffu, owned, last_lane);
    const Obs last_v = shfl_obs(v, last_lane);
    const bool last_first_nan = __shfl_sync(0xffffffffu, first_nan, last_lane);
    open = last_owned && !last_tail;
    carry = last_v;
    carry_first_nan = last_first_nan;
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
                   const uint32_t* counters, const LayerTable& lt, cudaStream_t s,
                   LaunchCounter& lc) {
  commit_move_clear_kernel<<<148 * 2, kBlock, 0, s>>>(p, st_in, st_out, counters, lt);
  ++lc.mine;
}
void launch_segreduce_estimate(const EstimateParams& p, const uint32_t* counters_ro,
                               uint32_t* counters, cudaStream_t s, LaunchCounter& lc) {
  if (p.n_sorted == 0) return;
  const size_t warps = (static_cast<size_t>(p.n_sorted) + kK3Chunk - 1) / kK3Chunk;
  const int grid = static_cast<int>((warps * 32 + kBlock - 1) / kBlock);
  segreduce_estimate_kernel<<<grid, kBlock, 0, s>>>(p, counters_ro, counters);
  ++lc.mine;
}
void launch_move_only(const DeviceState* st_in, DeviceState* st_out, double x, double y,
                      int clear_policy, const LayerTable& lt, uint32_t* moved_flag, cudaStream_t s,
                      LaunchCounter& lc) {
  move_only_kernel<<<148 * 2, kBlock, 0, s>>>(st_in, st_out, x, y, clear_policy, lt, moved_flag);
  ++lc.mine;
}

}  // namespace fdem
