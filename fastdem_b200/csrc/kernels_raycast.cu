// kernels_raycast.cu — BASELINE config 4's extras on the integrate() path:
//   voxelGrid(points, resolution, ANY)   nanopcl/filters/impl/voxel_grid_impl.hpp:30-236
//   applyRaycasting                      fastdem/src/raycasting.cpp:218-249
// plus the 3x3 inpainting stencil (fastdem/src/inpainting.cpp:21-67, a "next" row).
//
// Raycasting is not HBM-bound: it is a per-ray 2-D DDA whose cell visits are L2 atomics
// (atomicMin on an order-preserving encoding of the ray height) on a per-scan scratch
// buffer — scratch, never estimator state.  Compiled with -fmad=false; the DDA follows the
// reference's float32 arithmetic expression by expression.
#include <float.h>
#include <math.h>

#include "device_types.h"

namespace fdem {

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kEncInit = 0xffffffffu;   // "no ray crossed this cell"
constexpr uint32_t kNoSel = 0xffffffffu;
constexpr uint64_t kInvalidVoxel = ~0ull;    // nanopcl::voxel::INVALID_KEY (core/voxel.hpp:26)

__device__ __forceinline__ float nanf_() { return __int_as_float(0x7fc00000); }

// order-preserving float -> uint map (so atomicMin on the uint is a min on the float)
__device__ __forceinline__ uint32_t enc_f32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(uint32_t e) {
  const uint32_t u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  return __uint_as_float(u);
}

// voxel::pack (nanopcl/core/voxel.hpp:28-42): [z:21][y:21][x:21] of floor(p*inv) + 2^20
__device__ __forceinline__ uint64_t voxel_pack(float x, float y, float z, float inv) {
  constexpr int32_t OFF = 1 << 20;
  int32_t ix = static_cast<int32_t>(floorf(x * inv));
  int32_t iy = static_cast<int32_t>(floorf(y * inv));
  int32_t iz = static_cast<int32_t>(floorf(z * inv));
  ix = min(max(ix, -OFF), OFF - 1);
  iy = min(max(iy, -OFF), OFF - 1);
  iz = min(max(iz, -OFF), OFF - 1);
  return (static_cast<uint64_t>(iz + OFF) << 42) | (static_cast<uint64_t>(iy + OFF) << 21) |
         static_cast<uint64_t>(ix + OFF);
}

__global__ void __launch_bounds__(kBlock)
voxel_keys_kernel(const float4* __restrict__ pm, uint32_t n, float inv, uint64_t* __restrict__ keys,
                  uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = __ldg(&pm[i]);
  // dropped points are NaN-marked by K1; the reference skips non-finite points (:53-55)
  const bool ok = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
  keys[i] = ok ? voxel_pack(q.x, q.y, q.z, inv) : kInvalidVoxel;
  vals[i] = i;
}

// one representative per voxel: idx[start + (count*7 + start*13) % count]  (:171-172).
// The sort is stable, so within a voxel the indices ascend — the oracle's tie definition.
__global__ void __launch_bounds__(kBlock)
voxel_select_kernel(const uint64_t* __restrict__ skeys, const uint32_t* __restrict__ svals,
                    uint32_t n, uint32_t* __restrict__ counters, uint32_t* __restrict__ out_sel) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool head = false;
  if (i < n) {
    const uint64_t key = skeys[i];
    head = key != kInvalidVoxel && (i == 0 || skeys[i - 1] != key);
    uint32_t sel = kNoSel;
    if (head) {
      // upper bound of `key` in skeys[i, n)
      uint32_t lo = i + 1, hi = n;
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (skeys[mid] == key) lo = mid + 1; else hi = mid;
      }
      const uint64_t count = lo - i;
      const uint64_t start = i;
      sel = svals[start + (count * 7ull + start * 13ull) % count];
    }
    out_sel[i] = sel;
  }
  const uint32_t w = __popc(__ballot_sync(0xffffffffu, head));
  if ((threadIdx.x & 31) == 0 && w) atomicAdd(&counters[CNT_VOXELS], w);
}

// processScan (raycasting.cpp:150-179): one thread per ray_scan point
__global__ void __launch_bounds__(kBlock)
raycast_scan_kernel(const __grid_constant__ RaycastParams p, const DeviceState* __restrict__ st,
                    const float4* __restrict__ pts, const uint32_t* __restrict__ sel,
                    uint32_t n_max, uint32_t* __restrict__ counters) {
  const GridGeom g = st->geom;
  // preconditions (raycasting.cpp:230-234): sensor origin must be inside the map
  if (!geom_is_inside(g, static_cast<double>(p.origin[0]), static_cast<double>(p.origin[1]))) {
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CNT_RC_SKIP] = 1;
    return;
  }
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_max) return;
  uint32_t src = i;
  if (sel) {
    src = sel[i];
    if (src == kNoSel) return;
  }
  const float4 pt = __ldg(&pts[src]);
  const int nrows = g.rows, ncols = g.cols;

  // observed evidence: count hits per cell; the log-odds update itself is applied once
  // per hit, in order, by the resolve kernel (:162-170)
  {
    int32_t row, col;
    if (geom_get_index(g, static_cast<double>(pt.x), static_cast<double>(pt.y), row, col))
      atomicAdd(&p.hits[static_cast<size_t>(col) * nrows + row], 1u);
  }
  if (pt.z >= p.origin[2]) return;  // skip upward rays (:173)

  // traceRay (raycasting.cpp:46-139), float32 grid math
  const float resolution = static_cast<float>(g.res);
  const float sx = p.origin[0], sy = p.origin[1], sz = p.origin[2];
  const float dx = pt.x - sx;
  const float dy = pt.y - sy;
  const float ray_len_2d = sqrtf(dx * dx + dy * dy);
  if (ray_len_2d < 1e-4f) return;
  const float dz = pt.z - sz;
  const float origin_x = static_cast<float>(g.pos[0]) + nrows * resolution * 0.5f;
  const float origin_y = static_cast<float>(g.pos[1]) + ncols * resolution * 0.5f;
  const float gr0 = (origin_x - sx) / resolution;
  const float gc0 = (origin_y - sy) / resolution;
  const float gr1 = (origin_x - pt.x) / resolution;
  const float gc1 = (origin_y - pt.y) / resolution;
  const float dr = gr1 - gr0;
  const float dc = gc1 - gc0;
  int r = static_cast<int>(floorf(gr0));
  int c = static_cast<int>(floorf(gc0));
  int step_r, step_c;
  float t_max_r, t_max_c, t_delta_r, t_delta_c;
  if (fabsf(dr) > 1e-8f) {
    step_r = (dr > 0) ? 1 : -1;
    const float boundary = (step_r > 0) ? (r + 1.0f) : static_cast<float>(r);
    t_max_r = (boundary - gr0) / dr;
    t_delta_r = static_cast<float>(step_r) / dr;
  } else {
    step_r = 0;
    t_max_r = 1e30f;
    t_delta_r = 1e30f;
  }
  if (fabsf(dc) > 1e-8f) {
    step_c = (dc > 0) ? 1 : -1;
    const float boundary = (step_c > 0) ? (c + 1.0f) : static_cast<float>(c);
    t_max_c = (boundary - gc0) / dc;
    t_delta_c = static_cast<float>(step_c) / dc;
  } else {
    step_c = 0;
    t_max_c = 1e30f;
    t_delta_c = 1e30f;
  }
  const int max_steps = nrows + ncols;
  // The sensor cell is inside the map (precondition above) and the map is convex, so once
  // the ray has left the map it never comes back: the reference keeps stepping (its cells
  // fail the bounds test and are ignored); stopping there changes nothing but the work.
  bool was_inside = false;
  for (int s = 0; s < max_steps; ++s) {
    if (r >= 0 && r < nrows && c >= 0 && c < ncols) {
      was_inside = true;
      int mr = r + g.start[0];  // == (r + start) % size: both terms are in [0, size)
      if (mr >= nrows) mr -= nrows;
      int mc = c + g.start[1];
      if (mc >= ncols) mc -= ncols;
      const float t_exit = fminf(t_max_r, t_max_c);
      const float height = sz + fminf(t_exit, 1.0f) * dz;
      // Every ray starts in the sensor's cell, so the cells around it would take one atomic
      // per ray.  A minimum only ever decreases: if the value already stored (even a stale,
      // cached one — it can only be larger than the true current value) is <= mine, my
      // update cannot change anything and the atomic is skipped.
      uint32_t* slot = &p.ray_min_enc[static_cast<size_t>(mc) * nrows + mr];
      const uint32_t e = enc_f32(height);
      if (e < *slot) atomicMin(slot, e);  // plain (L1-cacheable) load: staleness is safe here
    } else if (was_inside) {
      break;
    }
    if (t_max_r < t_max_c) {
      if (t_max_r >= 1.0f) break;
      r += step_r;
      t_max_r += t_delta_r;
    } else {
      if (t_max_c >= 1.0f) break;
      c += step_c;
      t_max_c += t_delta_c;
    }
  }
}

// map.clear(raycasting) (:242) + the observed log-odds updates + resolveGhostCells
// (:188-214), one thread per cell; also rearms the scratch for the next scan.
__global__ void __launch_bounds__(kBlock)
raycast_resolve_kernel(const __grid_constant__ RaycastParams p,
                       const __grid_constant__ LayerTable lt,
                       const uint32_t* __restrict__ counters, size_t n_cells) {
  if (counters[CNT_RC_SKIP]) return;  // applyRaycasting returned before touching anything
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n_cells; i += stride) {
    const uint32_t hits = p.hits[i];
    const uint32_t enc = p.ray_min_enc[i];
    if (hits == 0 && enc == kEncInit) {
      p.raycasting[i] = nanf_();
      continue;
    }
    p.hits[i] = 0;
    p.ray_min_enc[i] = kEncInit;
    float lo = p.logodds[i];
    bool lo_dirty = false;
    if (hits) {
      if (isnan(lo)) lo = 0.0f;
      for (uint32_t k = 0; k < hits; ++k) {
        const float next = fminf(lo + p.log_odds_observed, p.log_odds_max);
        if (next == lo && p.log_odds_observed >= 0.0f) break;  // saturated: further hits are no-ops
        lo = next;
      }
      lo_dirty = true;
    }
    float rmin = nanf_();
    bool cleared = false;
    if (enc != kEncInit) {
      rmin = dec_f32(enc);
      const float elev = p.elevation[i];
      if (!isnan(elev) && elev > rmin + p.height_conflict_threshold) {
        if (isnan(lo)) lo = 0.0f;
        lo -= p.log_odds_ghost;
        lo_dirty = true;
        if (lo < p.clear_threshold) {
          // map.clearAt(idx): EVERY layer of the cell -> NaN, then ghost_removal = 1 (:208-211)
          for (int l = 0; l < lt.count; ++l) lt.ptr[l][i] = nanf_();
          p.ghost_removal[i] = 1.0f;
          cleared = true;
        }
      }
    }
    if (!cleared) {
      p.raycasting[i] = rmin;
      if (lo_dirty) p.logodds[i] = lo;
    }
  }
}

// applyInpainting, one Jacobi sweep: dst = src, NaN cells with >= min_valid finite 3x3
// logical neighbours become their mean (inpainting.cpp:41-62).  Neighbour order = the
// oracle's (dr outer, dc inner), so float sums agree exactly.
__global__ void __launch_bounds__(kBlock)
inpaint_iter_kernel(const float* __restrict__ src, float* __restrict__ dst,
                    const DeviceState* __restrict__ st, int min_valid, int rows_local, int cols) {
  const GridGeom g = st->geom;
  const size_t n = static_cast<size_t>(rows_local) * cols;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float v = src[i];
    float out = v;
    if (isnan(v)) {
      const int bc = static_cast<int>(i / rows_local);
      const int br = static_cast<int>(i - static_cast<size_t>(bc) * rows_local);
      const int lr = wrap_index(br - g.start[0] + g.rows, g.rows);
      const int lc = wrap_index(bc - g.start[1] + g.cols, g.cols);
      float sum = 0.0f;
      int count = 0;
      for (int dr = -1; dr <= 1; ++dr) {
        for (int dc = -1; dc <= 1; ++dc) {
          if (dr == 0 && dc == 0) continue;
          const int nr = lr + dr, nc = lc + dc;
          if (nr < 0 || nr >= g.rows || nc < 0 || nc >= g.cols) continue;
          const int nbr = wrap_index(nr + g.start[0], g.rows);
          const int nbc = wrap_index(nc + g.start[1], g.cols);
          const float val = src[static_cast<size_t>(nbc) * rows_local + nbr];
          if (isfinite(val)) {
            sum += val;
            ++count;
          }
        }
      }
      if (count >= min_valid) out = sum / static_cast<float>(count);
    }
    dst[i] = out;
  }
}

// applySpatialSmoothing (include/fastdem/postprocess/spatial_smoothing.hpp:38-67): median
// (rank size/2) of the finite values in the K x K logical neighbourhood, centre included.
template <int K>
__global__ void __launch_bounds__(kBlock)
median_filter_kernel(const float* __restrict__ src, float* __restrict__ dst,
                     const DeviceState* __restrict__ st, int min_valid, int rows_local, int cols) {
  const GridGeom g = st->geom;
  const size_t n = static_cast<size_t>(rows_local) * cols;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  constexpr int H = K / 2;
  for (; i < n; i += stride) {
    const float v = src[i];
    float out = v;
    if (isfinite(v)) {
      const int bc = static_cast<int>(i / rows_local);
      const int br = static_cast<int>(i - static_cast<size_t>(bc) * rows_local);
      const int lr = wrap_index(br - g.start[0] + g.rows, g.rows);
      const int lc = wrap_index(bc - g.start[1] + g.cols, g.cols);
      float w[K * K];
      int cnt = 0;
#pragma unroll
      for (int dr = -H; dr <= H; ++dr) {
#pragma unroll
        for (int dc = -H; dc <= H; ++dc) {
          const int nr = lr + dr, nc = lc + dc;
          if (nr < 0 || nr >= g.rows || nc < 0 || nc >= g.cols) continue;
          const float val = src[static_cast<size_t>(wrap_index(nc + g.start[1], g.cols)) * rows_local +
                                wrap_index(nr + g.start[0], g.rows)];
          if (isfinite(val)) {
            // insertion into the sorted prefix (tiny windows: 9 / 25 / 49 values)
            int j = cnt++;
            while (j > 0 && w[j - 1] > val) {
              w[j] = w[j - 1];
              --j;
            }
            w[j] = val;
          }
        }
      }
      if (cnt >= min_valid) out = w[cnt / 2];
    }
    dst[i] = out;
  }
}

inline int grid_for(size_t n, int block, int max_blocks = 148 * 8) {
  size_t b = (n + block - 1) / block;
  if (b < 1) b = 1;
  if (b > static_cast<size_t>(max_blocks)) b = max_blocks;
  return static_cast<int>(b);
}

}  // namespace

void launch_voxel_keys(const float4* pm, uint32_t n, float inv_voxel, uint64_t* keys,
                       uint32_t* vals, cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  voxel_keys_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(pm, n, inv_voxel, keys, vals);
  ++lc.mine;
}
void launch_voxel_select(const uint64_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                         uint32_t* counters, uint32_t* out_sel, cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  voxel_select_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(sorted_keys, sorted_vals, n,
                                                                  counters, out_sel);
  ++lc.mine;
}
void launch_raycast_scan(const RaycastParams& p, const DeviceState* st, const float4* pts,
                         const uint32_t* sel, const uint32_t* /*n_sel_dev*/, uint32_t n_max,
                         uint32_t* counters, cudaStream_t s, LaunchCounter& lc) {
  if (n_max == 0) return;
  raycast_scan_kernel<<<(n_max + kBlock - 1) / kBlock, kBlock, 0, s>>>(p, st, pts, sel, n_max,
                                                                      counters);
  ++lc.mine;
}
void launch_raycast_resolve(const RaycastParams& p, const DeviceState* /*st*/,
                            const LayerTable& lt, const uint32_t* counters, size_t n_cells,
                            cudaStream_t s, LaunchCounter& lc) {
  raycast_resolve_kernel<<<grid_for(n_cells, kBlock), kBlock, 0, s>>>(p, lt, counters, n_cells);
  ++lc.mine;
}
int launch_median_filter(const float* src, float* dst, const DeviceState* st, int kernel_size,
                         int min_valid, cudaStream_t s, LaunchCounter& lc, int rows_local, int cols) {
  const int grid = grid_for(static_cast<size_t>(rows_local) * cols, kBlock);
  switch (kernel_size) {
    case 1: median_filter_kernel<1><<<grid, kBlock, 0, s>>>(src, dst, st, min_valid, rows_local, cols); break;
    case 3: median_filter_kernel<3><<<grid, kBlock, 0, s>>>(src, dst, st, min_valid, rows_local, cols); break;
    case 5: median_filter_kernel<5><<<grid, kBlock, 0, s>>>(src, dst, st, min_valid, rows_local, cols); break;
    case 7: median_filter_kernel<7><<<grid, kBlock, 0, s>>>(src, dst, st, min_valid, rows_local, cols); break;
    default: return 1;
  }
  ++lc.mine;
  return 0;
}
void launch_inpaint_iter(const float* src, float* dst, const DeviceState* st, int min_valid,
                         cudaStream_t s, LaunchCounter& lc, int rows_local, int cols) {
  inpaint_iter_kernel<<<grid_for(static_cast<size_t>(rows_local) * cols, kBlock), kBlock, 0, s>>>(
      src, dst, st, min_valid, rows_local, cols);
  ++lc.mine;
}

}  // namespace fdem
