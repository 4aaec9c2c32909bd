// estimator.cuh — per-cell device code shared by the two cell-reduction paths
// (kernels.cu: global sort + K3; kernels_tile.cu: bucket partition + per-bucket K3t).
//
//   CellObs            one scan's observation of one cell, in a form whose combine() is
//                      associative AND commutative (explicit point-index tie-breaks), so any
//                      reduction order reproduces rasterize()'s sequential fold
//                      (fastdem/src/elevation_mapping.cpp:41-92)
//   kalman_cell / p2_cell   Kalman::update+computeBounds / P2Quantile::update+computeBounds
//   apply_observation  estimate() + updateMinMax/Obstacle/Intensity/Color for one cell
//
// -fmad=false: every expression is evaluated as written, in the oracle's order.
#pragma once

#include <float.h>
#include <math.h>

#include "device_types.h"

namespace fdem {

__device__ __forceinline__ float nan_f32() { return __int_as_float(0x7fc00000); }

// Programmatic dependent launch (PDL).  The scan's kernels form a chain in one CUDA graph
// with programmatic edges: each kernel lets its successor start launching right away
// (pdl_launch_dependents) and blocks at pdl_wait() until its predecessor has completed and
// flushed — so launch latency and prologues overlap the predecessor instead of adding up.
// Both are no-ops for a kernel launched without a programmatic dependency.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ElevationMapping::CellObservation (mapping/elevation_mapping.hpp:26-34) plus the point
// indices that decide rasterize()'s order-dependent choices:
//   min_z_var  = variance of the LOWEST-index point attaining min_z  (strict `z < min_z`, :65-68)
//   intensity  = first point's value taken blindly, then strict max  (:72-78)
//   color      = HIGHEST-index point of the cell                     (:81-88)
struct CellObs {
  float mz;      // min_z, FLT_MAX if no point entered through `z < min_z`
  float mv;      // min_z_var
  uint32_t mi;   // index of the point that set min_z (0xffffffff = none)
  float xz;      // max_z, -FLT_MAX if none
  float it;      // max over non-NaN intensities, -inf if none
  uint32_t fi;   // (lowest point index << 1) | (that point's intensity is NaN)
  uint32_t li;   // highest point index
};

__device__ __forceinline__ CellObs obs_identity() {
  CellObs o;
  o.mz = FLT_MAX; o.mv = 0.0f; o.mi = 0xffffffffu; o.xz = -FLT_MAX; o.it = -INFINITY;
  o.fi = 0xffffffffu; o.li = 0u;
  return o;
}

// one point as an observation
__device__ __forceinline__ CellObs obs_from_point(float z, float var_z, float intensity,
                                                  bool has_intensity, uint32_t idx) {
  CellObs o = obs_identity();
  // a point only enters through a strict compare against the initial values, so NaN /
  // out-of-range z leave them in place (edge case 4 of SURVEY.md Appendix C)
  if (z < FLT_MAX) { o.mz = z; o.mv = var_z; o.mi = idx; }
  if (z > -FLT_MAX) o.xz = z;
  bool nan_i = false;
  if (has_intensity) {
    nan_i = isnan(intensity);
    if (!nan_i) o.it = intensity;
  }
  o.fi = (idx << 1) | (nan_i ? 1u : 0u);
  o.li = idx;
  return o;
}

__device__ __forceinline__ CellObs obs_combine(const CellObs& a, const CellObs& b) {
  CellObs r;
  const bool take_b = (b.mz < a.mz) || (b.mz == a.mz && b.mi < a.mi);
  r.mz = take_b ? b.mz : a.mz;
  r.mv = take_b ? b.mv : a.mv;
  r.mi = take_b ? b.mi : a.mi;
  r.xz = (b.xz > a.xz) ? b.xz : a.xz;
  r.it = (b.it > a.it) ? b.it : a.it;
  r.fi = min(a.fi, b.fi);
  r.li = max(a.li, b.li);
  return r;
}

__device__ __forceinline__ CellObs obs_shfl_up(const CellObs& v, int d) {
  CellObs r;
  r.mz = __shfl_up_sync(0xffffffffu, v.mz, d);
  r.mv = __shfl_up_sync(0xffffffffu, v.mv, d);
  r.mi = __shfl_up_sync(0xffffffffu, v.mi, d);
  r.xz = __shfl_up_sync(0xffffffffu, v.xz, d);
  r.it = __shfl_up_sync(0xffffffffu, v.it, d);
  r.fi = __shfl_up_sync(0xffffffffu, v.fi, d);
  r.li = __shfl_up_sync(0xffffffffu, v.li, d);
  return r;
}
__device__ __forceinline__ CellObs obs_shfl(const CellObs& v, int src) {
  CellObs r;
  r.mz = __shfl_sync(0xffffffffu, v.mz, src);
  r.mv = __shfl_sync(0xffffffffu, v.mv, src);
  r.mi = __shfl_sync(0xffffffffu, v.mi, src);
  r.xz = __shfl_sync(0xffffffffu, v.xz, src);
  r.it = __shfl_sync(0xffffffffu, v.it, src);
  r.fi = __shfl_sync(0xffffffffu, v.fi, src);
  r.li = __shfl_sync(0xffffffffu, v.li, src);
  return r;
}

// Per-cell state is loaded in ONE batch before any arithmetic or store: the loads are
// independent, so a touched cell costs one memory round trip instead of one per layer
// (stores to one layer would otherwise fence the compiler from hoisting the next load).
struct CellState {
  float smin, smax, sint;   // elevation_min, elevation_max, intensity
  uint32_t rgb_bits;        // packed colour of the cell's last point
  // Kalman
  float x, P, count, mean, svar, m2;
  // P2
  float q[5], n[5];
};

__device__ __forceinline__ CellState load_cell_state(const EstimateParams& p, uint32_t c,
                                                     const CellObs& v) {
  const EstLayers& L = p.L;
  CellState s;
  s.smin = L.elevation_min[c];
  s.smax = L.elevation_max[c];
  s.count = L.n_points[c];
  s.sint = p.intensity ? L.intensity[c] : 0.0f;
  s.rgb_bits = 0;
  if (p.rgb) {
    const uint8_t* rgb = p.rgb + static_cast<size_t>(v.li) * 3;
    s.rgb_bits = (static_cast<uint32_t>(rgb[0]) << 16) | (static_cast<uint32_t>(rgb[1]) << 8) | rgb[2];
  }
  if (p.estimation_type == 1) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      s.q[k] = L.p2_q[k][c];
      s.n[k] = L.p2_n[k][c];
    }
    s.x = s.P = s.mean = s.svar = s.m2 = 0.0f;
  } else {
    s.x = L.elevation[c];
    s.P = L.kalman_p[c];
    s.mean = L.sample_mean[c];
    s.svar = L.variance[c];
    s.m2 = L.sample_m2[c];
#pragma unroll
    for (int k = 0; k < 5; ++k) s.q[k] = s.n[k] = 0.0f;
  }
  return s;
}

// Kalman::update + computeBounds on one cell (mapping/kalman_estimation.hpp:98-153)
__device__ __forceinline__ void kalman_cell(const EstimateParams& p, uint32_t c, CellState& s,
                                            float z, float meas_var) {
  const EstLayers& L = p.L;
  float x = s.x, P = s.P, count = s.count, mean = s.mean, svar = s.svar, m2 = s.m2;
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
    // std::clamp(P, min, max) exactly (kalman_estimation.hpp:124): a NaN P stays NaN
    P = (P < p.kalman_min_variance) ? p.kalman_min_variance
                                    : ((p.kalman_max_variance < P) ? p.kalman_max_variance : P);
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
__device__ __forceinline__ void p2_cell(const EstimateParams& p, uint32_t c, CellState& s,
                                        float x) {
  const EstLayers& L = p.L;
  float q[5], n[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    q[k] = s.q[k];
    n[k] = s.n[k];
  }
  float count = s.count;
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

// Everything ElevationMapping::update does to ONE touched cell `c` given the scan's
// observation of it and the cell's pre-loaded state: estimate() (:94-108), updateMinMax
// (:127-142), updateObstacle (:144-152), updateIntensity (:154-166), updateColor (:168-175).
// Called exactly once per touched cell per scan; every layer value is a plain store.
__device__ __forceinline__ void apply_observation(const EstimateParams& p, uint32_t c,
                                                  const CellObs& v, CellState& s) {
  if (p.estimation_type == 1) p2_cell(p, c, s, v.mz);
  else kalman_cell(p, c, s, v.mz, v.mv);
  if (isnan(s.smin) || v.mz < s.smin) p.L.elevation_min[c] = v.mz;
  if (isnan(s.smax) || v.xz > s.smax) p.L.elevation_max[c] = v.xz;
  p.L.obstacle[c] = (v.xz > v.mz) ? v.xz : nan_f32();
  if (p.intensity) {
    // the per-scan max is NaN only when the cell's first point carries NaN (rasterize
    // takes the first value blindly, :72-78)
    const float mi = (v.fi & 1u) ? nan_f32() : v.it;
    if (isnan(s.sint) || mi > s.sint) p.L.intensity[c] = mi;
  }
  // last point of the cell wins; 0x00RRGGBB reinterpreted as float (colorVectorToValue)
  if (p.rgb) reinterpret_cast<uint32_t*>(p.L.color)[c] = s.rgb_bits;
}

__device__ __forceinline__ void apply_observation(const EstimateParams& p, uint32_t c,
                                                  const CellObs& v) {
  CellState s = load_cell_state(p, c, v);
  apply_observation(p, c, v, s);
}

}  // namespace fdem
