// kernels_raycast.cu — BASELINE config 4's extras on the integrate() path:
//   voxelGrid(points, resolution, ANY)   nanopcl/filters/impl/voxel_grid_impl.hpp:30-236
//   applyRaycasting                      fastdem/src/raycasting.cpp:218-249
// plus the 3x3 inpainting stencil (fastdem/src/inpainting.cpp:21-67, a "next" row).
//
// Raycasting is not HBM-bound: it is a per-ray 2-D DDA whose cell visits are reads of (and rare
// atomics on) a per-scan, L2-resident scratch buffer — scratch, never estimator state.  Compiled with -fmad=false; the DDA follows the
// reference's float32 arithmetic expression by expression.
#include <float.h>
#include <cstdlib>
#include <math.h>

#include "device_types.h"

namespace fdem {

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kNoSel = 0xffffffffu;
constexpr uint64_t kInvalidVoxel = ~0ull;    // nanopcl::voxel::INVALID_KEY (core/voxel.hpp:26)

__device__ __forceinline__ float nanf_() { return __int_as_float(0x7fc00000); }

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

// Compact voxel keys.  When the crop filters bound the kept points to a box the host can
// compute (finite range_max; see voxel_box() in capi.cu), the same [z][y][x] order fits a
// 32-bit key: (iz - z0) << (bx+by) | (iy - y0) << bx | (ix - x0).  Subtracting a constant per
// axis preserves the lexicographic order of voxel::pack, so the sorted sequence — and with
// it every voxel's `start` rank — is identical to the 63-bit key's; the radix sort then
// needs 4 passes over 8-byte pairs instead of 8 passes over 12-byte pairs.
__global__ void __launch_bounds__(kBlock)
voxel_keys32_kernel(const float4* __restrict__ pm, uint32_t n, float inv, const VoxelBox box,
                    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                    uint32_t* __restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = __ldg(&pm[i]);
  const bool ok = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
  uint32_t key = box.invalid_key;
  if (ok) {
    constexpr int32_t OFF = 1 << 20;
    int32_t ix = static_cast<int32_t>(floorf(q.x * inv));
    int32_t iy = static_cast<int32_t>(floorf(q.y * inv));
    int32_t iz = static_cast<int32_t>(floorf(q.z * inv));
    ix = min(max(ix, -OFF), OFF - 1) - box.x0;
    iy = min(max(iy, -OFF), OFF - 1) - box.y0;
    iz = min(max(iz, -OFF), OFF - 1) - box.z0;
    const int32_t mx = (1 << box.bx) - 1, my = (1 << box.by) - 1, mz = (1 << box.bz) - 1;
    if (ix < 0 || ix > mx || iy < 0 || iy > my || iz < 0 || iz > mz) {
      // cannot happen while the transforms are rigid; counted so a test / caller can see it
      atomicAdd(&counters[CNT_VOX_VIOLATION], 1u);
      ix = min(max(ix, 0), mx); iy = min(max(iy, 0), my); iz = min(max(iz, 0), mz);
    }
    key = (static_cast<uint32_t>(iz) << (box.bx + box.by)) | (static_cast<uint32_t>(iy) << box.bx) |
          static_cast<uint32_t>(ix);
  }
  keys[i] = key;
  vals[i] = i;
}

// one representative per voxel: idx[start + (count*7 + start*13) % count]  (:171-172).
// The sort is stable, so within a voxel the indices ascend — the oracle's tie definition.
template <typename Key>
__global__ void __launch_bounds__(kBlock)
voxel_select_kernel(const Key* __restrict__ skeys, const uint32_t* __restrict__ svals, uint32_t n,
                    const Key invalid, uint32_t* __restrict__ counters,
                    uint32_t* __restrict__ out_sel) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool head = false;
  if (i < n) {
    const Key key = skeys[i];
    head = key != invalid && (i == 0 || skeys[i - 1] != key);
    uint32_t sel = kNoSel;
    if (head) {
      // end of the run of `key`: voxels hold a point or two, so walk a few elements first
      // (coalesced with the neighbours' walks) and binary-search only a long run
      uint32_t lo = i + 1;
      constexpr uint32_t kWalk = 6;
      const uint32_t wend = min(n, i + 1 + kWalk);
      while (lo < wend && skeys[lo] == key) ++lo;
      if (lo == wend && lo < n && skeys[lo] == key) {
        uint32_t hi = n;
        ++lo;
        while (lo < hi) {
          const uint32_t mid = lo + ((hi - lo) >> 1);
          if (skeys[mid] == key) lo = mid + 1; else hi = mid;
        }
      }
      const uint64_t count = lo - i;
      const uint64_t start = i;
      // (count*7 + start*13) % count; a voxel with one point needs no 64-bit division
      sel = svals[count == 1 ? start : start + (count * 7ull + start * 13ull) % count];
    }
    out_sel[i] = sel;
  }
  // one atomic per CTA: thousands of same-address atomics would serialise in L2
  const int heads = __syncthreads_count(head);
  if (threadIdx.x == 0 && heads) atomicAdd(&counters[CNT_VOXELS], static_cast<uint32_t>(heads));
}

// ── processScan's per-point part, fused with the voxel-representative selection ──
// One pass over the voxel-sorted (key, index) stream does everything raycasting.cpp:160-174
// does per ray_scan point EXCEPT the DDA itself:
//   * voxelGrid(ANY) representative of every voxel (voxel_grid_impl.hpp:171-172)
//   * observed evidence: one hit count per representative that falls inside the map (the
//     log-odds additions themselves are applied per hit, in order, by the resolve kernel)
//   * the rays that will be traced (downward, >= 1e-4 m long in xy) are appended to a list of
//     end points together with a 15-bit ordering key (length bin, azimuth bin) whose histogram
//     is built on the fly — the counting sort below turns the list into BUNDLES: 32 consecutive
//     rays of the sorted list point the same way (< 1 degree apart) and are equally long, so a
//     warp of the DDA kernel walks 32 nearly identical rays (coalesced loads, lanes that finish
//     together).  Order has no effect on the result: a cell keeps the MINIMUM over its rays.
// skeys == nullptr: applyRaycasting on a caller's cloud (fdem_raycast) — every point is a ray_scan point.
constexpr int kRayLenBins = 32;    // ray length in 32-cell bins
constexpr int kRayAzBins = 1024;   // "diamond angle" sectors (monotone in azimuth, no atan2)
constexpr int kRayBins = kRayLenBins * kRayAzBins;

template <typename Key>
__global__ void __launch_bounds__(kBlock)
voxel_select_rays_kernel(const Key* __restrict__ skeys, const uint32_t* __restrict__ svals,
                         uint32_t n, const Key invalid, const __grid_constant__ RaycastParams p,
                         const DeviceState* __restrict__ st, const float4* __restrict__ pts,
                         uint32_t* __restrict__ counters, float4* __restrict__ rays_unsorted,
                         uint32_t* __restrict__ ray_hist) {
  __shared__ uint32_t s_warp[kBlock / 32];
  __shared__ uint32_t s_base;
  __shared__ uint32_t s_len[kRayLenBins];   // rays per length bin of this block (one global atomic per bin)
  const GridGeom g = st->geom;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kRayLenBins) s_len[threadIdx.x] = 0u;
  __syncthreads();
  // preconditions (raycasting.cpp:230-234): sensor origin must be inside the map; the voxel
  // count is still reported (voxelGrid runs before applyRaycasting, fastdem.cpp:156-158)
  const bool rc_ok = geom_is_inside(g, static_cast<double>(p.origin[0]), static_cast<double>(p.origin[1]));
  if (!rc_ok && i == 0) counters[CNT_RC_SKIP] = 1;
  bool head = false, trace = false;
  float4 pt = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (i < n) {
    uint32_t src = kNoSel;
    if (skeys) {
      const Key key = skeys[i];
      head = key != invalid && (i == 0 || skeys[i - 1] != key);
      if (head) {
        // end of the run of `key`: voxels hold a point or two, so walk a few elements first
        // (coalesced with the neighbours' walks) and binary-search only a long run
        uint32_t lo = i + 1;
        constexpr uint32_t kWalk = 6;
        const uint32_t wend = min(n, i + 1 + kWalk);
        while (lo < wend && skeys[lo] == key) ++lo;
        if (lo == wend && lo < n && skeys[lo] == key) {
          uint32_t hi = n;
          ++lo;
          while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (skeys[mid] == key) lo = mid + 1; else hi = mid;
          }
        }
        const uint64_t count = lo - i;
        const uint64_t start = i;
        src = svals[count == 1 ? start : start + (count * 7ull + start * 13ull) % count];
      }
    } else {
      head = true;
      src = i;
    }
    if (head && rc_ok) {
      pt = __ldg(&pts[src]);
      int32_t row, col;
      if (geom_get_index(g, static_cast<double>(pt.x), static_cast<double>(pt.y), row, col))
        atomicAdd(&p.hits[static_cast<size_t>(col) * g.rows + row], 1u);
      const float dx = pt.x - p.origin[0];
      const float dy = pt.y - p.origin[1];
      const float ray_len_2d = sqrtf(dx * dx + dy * dy);
      // upward rays are skipped (:173); rays shorter than 1e-4 m in xy are skipped (:52-53)
      trace = pt.z < p.origin[2] && ray_len_2d >= 1e-4f;
      if (trace) {
        // ordering key (grouping only, no exactness needed): DDA steps ~ (|dx| + |dy|) / res
        const float l1 = fabsf(dx) + fabsf(dy);
        const int lenbin = min(static_cast<int>(l1 / static_cast<float>(g.res)) >> 5, kRayLenBins - 1);
        float a = dy / l1;                        // [-1, 1]
        if (dx < 0.0f) a = 2.0f - a;              // (1, 3]
        else if (dy < 0.0f) a = 4.0f + a;         // [3, 4)
        const int azbin = (min(max(static_cast<int>(a * (kRayAzBins / 4)), 0), kRayAzBins - 1)) & p.tune_az_mask;
        const uint32_t okey = static_cast<uint32_t>(lenbin * kRayAzBins + azbin);
        pt.w = __uint_as_float(okey);
        atomicAdd(&ray_hist[okey], 1u);
        atomicAdd(&s_len[lenbin], 1u);   // rays per length bin (segment bases)
      }
    }
  }
  // compaction inside the block, one atomic per block for its base (scan scratch)
  const uint32_t tm = __ballot_sync(0xffffffffu, trace);
  const uint32_t hm = __ballot_sync(0xffffffffu, head);
  if (lane == 0) s_warp[warp] = (static_cast<uint32_t>(__popc(tm)) << 16) | static_cast<uint32_t>(__popc(hm));
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t rays_total = 0, heads_total = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) {
      const uint32_t v = s_warp[w];
      s_warp[w] = rays_total;           // exclusive prefix of the traced rays
      rays_total += v >> 16;
      heads_total += v & 0xffffu;
    }
    if (skeys && heads_total) atomicAdd(&counters[CNT_VOXELS], heads_total);
    s_base = rays_total ? atomicAdd(&counters[CNT_RAYS], rays_total) : 0u;
  }
  if (threadIdx.x < kRayLenBins && s_len[threadIdx.x]) atomicAdd(&ray_hist[2 * kRayBins + threadIdx.x], s_len[threadIdx.x]);
  __syncthreads();
  if (trace) rays_unsorted[s_base + s_warp[warp] + __popc(tm & ((1u << lane) - 1u))] = pt;
}

// counting sort of the ray list by ordering key, step 2 of 3: one CTA per length bin — base of
// the bin's segment from the per-length-bin totals, exclusive scan over its azimuth bins,
// cursors out, histogram re-armed for the next scan
__global__ void __launch_bounds__(kRayAzBins)
ray_bin_scan_kernel(uint32_t* __restrict__ ray_hist, uint32_t* __restrict__ ray_cursor) {
  __shared__ uint32_t s_warp[kRayAzBins / 32];
  __shared__ uint32_t s_base;
  const int L = blockIdx.x, t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const uint32_t* len_hist = ray_hist + 2 * kRayBins;
  const uint32_t v = ray_hist[L * kRayAzBins + t];
  ray_hist[L * kRayAzBins + t] = 0u;
  if (warp == 0) {
    uint32_t x = (lane < L) ? len_hist[lane] : 0u;   // kRayLenBins == 32: one lane per length bin
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if (lane == 0) s_base = x;
  }
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t wprefix = 0;
  for (int w = 0; w < warp; ++w) wprefix += s_warp[w];
  ray_cursor[L * kRayAzBins + t] = s_base + wprefix + inc - v;
}

// step 3 of 3: scatter (order inside a bin is irrelevant)
__global__ void __launch_bounds__(kBlock)
ray_bin_scatter_kernel(const float4* __restrict__ rays_unsorted, float4* __restrict__ rays,
                       uint32_t* __restrict__ ray_cursor, const uint32_t* __restrict__ counters) {
  const uint32_t n_rays = counters[CNT_RAYS];
  // the per-length-bin totals were consumed by the scan kernel: re-arm them
  if (blockIdx.x == 0 && threadIdx.x < kRayLenBins) ray_cursor[kRayBins + threadIdx.x] = 0u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += gridDim.x * blockDim.x) {
    const float4 pt = rays_unsorted[i];
    rays[atomicAdd(&ray_cursor[__float_as_uint(pt.w)], 1u)] = pt;
  }
}

// ═════════════ voxelGrid(ANY) without a library sort: a 2-level MSD sort built for it ═════════════
// The voxel key is [z][y][x]; what voxelGrid needs from the sort is, per voxel, its point count,
// its global rank `start` among the valid points, and its points in index order
// (voxel_grid_impl.hpp:63,171-172).  Level 1 partitions the points by ROW = (z, y) — a 5 cm x
// 5 cm strip along x — with one histogram pass, one scan over the rows and one scatter; rows are
// in key order, so a row's base in the scan IS the global rank of its first point.  Level 2 sorts
// every row by (x, index) on chip: a row of <= 32 points in the registers of one warp (bitonic
// over shuffles, no barrier), a bigger row in the shared memory of one CTA (rows beyond its
// capacity: the same network in global memory — slow but exact).  The representative choice and
// everything after it (hits, ray list) is fused behind the sort exactly as in
// voxel_select_rays_kernel.  ~6 short kernels of our own instead of 7 library launches moving
// 8-byte pairs four times through HBM.
constexpr int kRowBlock = 1024;          // rows scanned per CTA of the row-scan kernel
constexpr int kBigRowSmem = 4096;        // pairs a CTA sorts in shared memory

struct VoxelRows {
  uint32_t* row_count;    // [n_rows] points per row; all zero between scans
  uint32_t* row_excl;     // [n_rows] exclusive prefix inside the row's block of kRowBlock rows
  uint32_t* row_cursor;   // [n_rows] scatter cursors
  uint32_t* block_sum;    // [n_blocks]
  uint32_t* block_base;   // [n_blocks] exclusive prefix over the blocks
  uint32_t* small_list;   // [n_rows] rows with 1..32 points
  uint32_t* big_list;     // [n_rows] rows with more
  uint32_t* ctl;          // [8]: 0 n_small, 1 n_big, 2 ticket, 3 small work, 4 big work
  uint32_t* vkey;         // [n] voxel key per point (invalid_key: dropped)
  unsigned long long* pairs;  // [n] (x << 32 | point index), grouped by row
  uint32_t n_rows, n_blocks, bx, invalid_key;
};

// level 1a: key per point + rows histogram.  Consecutive points of a scan line often fall into
// the same row: one atomic per run of equal rows in a warp.
__global__ void __launch_bounds__(kBlock)
voxel_rows_kernel(const float4* __restrict__ pm, uint32_t n, float inv, const VoxelBox box,
                  const __grid_constant__ VoxelRows vr, uint32_t* __restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (i < 8) vr.ctl[i] = 0u;   // lists and work counters of this scan (the previous scan's consumers are done)
  uint32_t key = box.invalid_key;
  if (i < n) {
    const float4 q = __ldg(&pm[i]);
    if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z)) {
      constexpr int32_t OFF = 1 << 20;
      int32_t ix = static_cast<int32_t>(floorf(q.x * inv));
      int32_t iy = static_cast<int32_t>(floorf(q.y * inv));
      int32_t iz = static_cast<int32_t>(floorf(q.z * inv));
      ix = min(max(ix, -OFF), OFF - 1) - box.x0;
      iy = min(max(iy, -OFF), OFF - 1) - box.y0;
      iz = min(max(iz, -OFF), OFF - 1) - box.z0;
      const int32_t mx = (1 << box.bx) - 1, my = (1 << box.by) - 1, mz = (1 << box.bz) - 1;
      if (ix < 0 || ix > mx || iy < 0 || iy > my || iz < 0 || iz > mz) {
        atomicAdd(&counters[CNT_VOX_VIOLATION], 1u);   // cannot happen while the transforms are rigid
        ix = min(max(ix, 0), mx); iy = min(max(iy, 0), my); iz = min(max(iz, 0), mz);
      }
      key = (static_cast<uint32_t>(iz) << (box.bx + box.by)) | (static_cast<uint32_t>(iy) << box.bx) |
            static_cast<uint32_t>(ix);
    }
    vr.vkey[i] = key;
  }
  const uint32_t row = key == box.invalid_key ? 0xffffffffu : key >> vr.bx;
  const uint32_t prev = __shfl_up_sync(0xffffffffu, row, 1);
  const bool head = lane == 0 || row != prev;
  const uint32_t hm = __ballot_sync(0xffffffffu, head);
  if (head && row != 0xffffffffu) {
    const uint32_t above = hm & ~((2u << lane) - 1u);           // heads after mine
    const int end = above ? __ffs(above) - 1 : 32;
    atomicAdd(&vr.row_count[row], static_cast<uint32_t>(end - lane));
  }
}

// level 1b: exclusive scan over the rows (per block of kRowBlock rows + the last block to finish
// scans the block totals), the two work lists, and the re-arming of the histogram
__global__ void __launch_bounds__(kBlock)
voxel_row_scan_kernel(const __grid_constant__ VoxelRows vr) {
  __shared__ uint32_t s_warp[kBlock / 32];
  __shared__ uint32_t s_small[kRowBlock], s_big[kRowBlock];
  __shared__ uint32_t s_ns, s_nb, s_bs, s_bb, s_last;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) { s_ns = 0; s_nb = 0; }
  __syncthreads();
  constexpr int PER = kRowBlock / kBlock;   // 4 consecutive rows per thread
  const uint32_t r0 = blockIdx.x * kRowBlock + t * PER;
  uint32_t c[PER];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    c[k] = (r0 + k < vr.n_rows) ? vr.row_count[r0 + k] : 0u;
    sum += c[k];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t wprefix = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kBlock / 32; ++w) {
    if (w < warp) wprefix += s_warp[w];
    total += s_warp[w];
  }
  uint32_t base = wprefix + inc - sum;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    if (r0 + k < vr.n_rows) {
      vr.row_excl[r0 + k] = base;
      vr.row_cursor[r0 + k] = 0u;
      if (c[k]) {
        vr.row_count[r0 + k] = 0u;   // re-arm (rows that stayed empty are zero already)
        if (c[k] <= 32u) s_small[atomicAdd(&s_ns, 1u)] = r0 + k;
        else s_big[atomicAdd(&s_nb, 1u)] = r0 + k;
      }
    }
    base += c[k];
  }
  __syncthreads();
  if (t == 0) {
    s_bs = s_ns ? atomicAdd(&vr.ctl[0], s_ns) : 0u;
    s_bb = s_nb ? atomicAdd(&vr.ctl[1], s_nb) : 0u;
    vr.block_sum[blockIdx.x] = total;
    __threadfence();
    s_last = (atomicAdd(&vr.ctl[2], 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  // lists keep (row, count): the consumers need both
  for (uint32_t k = t; k < s_ns; k += kBlock) vr.small_list[s_bs + k] = s_small[k];
  for (uint32_t k = t; k < s_nb; k += kBlock) vr.big_list[s_bb + k] = s_big[k];
  if (s_last) {
    // exclusive scan of the block totals (n_blocks <= a few thousand): this CTA, PER per thread in rounds
    __threadfence();
    __shared__ uint32_t s_run;
    if (t == 0) s_run = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < vr.n_blocks; b0 += kBlock) {
      const uint32_t b = b0 + t;
      const uint32_t v = b < vr.n_blocks ? __ldcg(&vr.block_sum[b]) : 0u;
      uint32_t in2 = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, in2, d);
        if (lane >= d) in2 += o;
      }
      if (lane == 31) s_warp[warp] = in2;
      __syncthreads();
      uint32_t wp = 0, all = 0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; ++w) {
        if (w < warp) wp += s_warp[w];
        all += s_warp[w];
      }
      if (b < vr.n_blocks) vr.block_base[b] = s_run + wp + in2 - v;
      __syncthreads();
      if (t == 0) s_run += all;
      __syncthreads();
    }
  }
}

// level 1c: scatter (x, index) into the row segments
__global__ void __launch_bounds__(kBlock)
voxel_scatter_kernel(uint32_t n, const __grid_constant__ VoxelRows vr) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t key = i < n ? vr.vkey[i] : vr.invalid_key;
  const uint32_t row = key == vr.invalid_key ? 0xffffffffu : key >> vr.bx;
  const uint32_t prev = __shfl_up_sync(0xffffffffu, row, 1);
  const bool head = lane == 0 || row != prev;
  const uint32_t hm = __ballot_sync(0xffffffffu, head);
  const int my_head = 31 - __clz(hm & (0xffffffffu >> (31 - lane)));   // lane 0 is always a head
  uint32_t base = 0;
  if (head && row != 0xffffffffu) {
    const uint32_t above = hm & ~((2u << lane) - 1u);
    const int end = above ? __ffs(above) - 1 : 32;
    base = vr.block_base[row / kRowBlock] + vr.row_excl[row] +
           atomicAdd(&vr.row_cursor[row], static_cast<uint32_t>(end - lane));
  }
  base = __shfl_sync(0xffffffffu, base, my_head);
  if (row != 0xffffffffu)
    vr.pairs[base + (lane - my_head)] =
        (static_cast<unsigned long long>(key & ((1u << vr.bx) - 1u)) << 32) | i;
}

// what follows the sort for one voxel representative (processScan's per-point part,
// raycasting.cpp:160-174): hit count, ray test, ordering key; returns whether the ray is traced
__device__ __forceinline__ bool ray_of_representative(const RaycastParams& p, const GridGeom& g, float4& pt,
                                                      uint32_t* __restrict__ ray_hist, uint32_t* s_len) {
  int32_t row, col;
  if (geom_get_index(g, static_cast<double>(pt.x), static_cast<double>(pt.y), row, col))
    atomicAdd(&p.hits[static_cast<size_t>(col) * g.rows + row], 1u);
  const float dx = pt.x - p.origin[0];
  const float dy = pt.y - p.origin[1];
  const float ray_len_2d = sqrtf(dx * dx + dy * dy);
  if (!(pt.z < p.origin[2] && ray_len_2d >= 1e-4f)) return false;   // upward / degenerate rays (:173, :52-53)
  const float l1 = fabsf(dx) + fabsf(dy);
  const int lenbin = min(static_cast<int>(l1 / static_cast<float>(g.res)) >> 5, kRayLenBins - 1);
  float a = dy / l1;
  if (dx < 0.0f) a = 2.0f - a;
  else if (dy < 0.0f) a = 4.0f + a;
  const int azbin = (min(max(static_cast<int>(a * (kRayAzBins / 4)), 0), kRayAzBins - 1)) & p.tune_az_mask;
  const uint32_t okey = static_cast<uint32_t>(lenbin * kRayAzBins + azbin);
  pt.w = __uint_as_float(okey);
  atomicAdd(&ray_hist[okey], 1u);
  atomicAdd(&s_len[lenbin], 1u);
  return true;
}

// all-ascending bitonic network: compare-exchange (i, l) always leaves the minimum at the lower
// index, so virtual +inf padding above the data never moves
__device__ __forceinline__ unsigned long long warp_bitonic32(unsigned long long v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
    {
      const int l = lane ^ (k - 1);
      const unsigned long long o = __shfl_sync(0xffffffffu, v, l);
      v = (lane < l) ? (o < v ? o : v) : (o > v ? o : v);
    }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
      const int l = lane ^ j;
      const unsigned long long o = __shfl_sync(0xffffffffu, v, l);
      v = (lane < l) ? (o < v ? o : v) : (o > v ? o : v);
    }
  }
  return v;
}

// level 2, rows of <= 32 points: one warp per row, sorted in registers
__global__ void __launch_bounds__(kBlock)
voxel_rows_small_kernel(const __grid_constant__ VoxelRows vr, const __grid_constant__ RaycastParams p,
                        const DeviceState* __restrict__ st, const float4* __restrict__ pts,
                        uint32_t* __restrict__ counters, float4* __restrict__ rays_unsorted,
                        uint32_t* __restrict__ ray_hist) {
  __shared__ uint32_t s_len[kRayLenBins];
  if (threadIdx.x < kRayLenBins) s_len[threadIdx.x] = 0u;
  __syncthreads();
  const GridGeom g = st->geom;
  const int lane = threadIdx.x & 31;
  const bool rc_ok = geom_is_inside(g, static_cast<double>(p.origin[0]), static_cast<double>(p.origin[1]));
  if (!rc_ok && blockIdx.x == 0 && threadIdx.x == 0) counters[CNT_RC_SKIP] = 1;
  const uint32_t n_jobs = vr.ctl[0];
  uint32_t voxels = 0;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t job = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; job < n_jobs; job += warps) {
    const uint32_t row = vr.small_list[job];
    const uint32_t base = vr.block_base[row / kRowBlock] + vr.row_excl[row];
    const uint32_t c = vr.row_cursor[row];   // == the row's point count after the scatter
    unsigned long long v = lane < static_cast<int>(c) ? vr.pairs[base + lane] : ~0ull;
    v = warp_bitonic32(v, lane);
    const uint32_t x = static_cast<uint32_t>(v >> 32);
    const uint32_t xp = __shfl_up_sync(0xffffffffu, x, 1);
    const bool valid = lane < static_cast<int>(c);
    const bool head = valid && (lane == 0 || x != xp);
    const uint32_t hm = __ballot_sync(0xffffffffu, head);
    int sel = lane;
    if (head) {
      const uint32_t above = hm & ~((2u << lane) - 1u);
      const uint64_t count = static_cast<uint64_t>((above ? __ffs(above) - 1 : static_cast<int>(c)) - lane);
      const uint64_t start = static_cast<uint64_t>(base) + lane;   // global rank among the valid points
      sel = lane + static_cast<int>(count == 1 ? 0ull : (count * 7ull + start * 13ull) % count);
    }
    const uint32_t src = static_cast<uint32_t>(__shfl_sync(0xffffffffu, v, sel) & 0xffffffffull);
    voxels += __popc(hm);
    bool trace = false;
    float4 pt = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (head && rc_ok) {
      pt = __ldg(&pts[src]);
      trace = ray_of_representative(p, g, pt, ray_hist, s_len);
    }
    const uint32_t tm = __ballot_sync(0xffffffffu, trace);
    if (tm) {
      uint32_t rb = 0;
      if (lane == 0) rb = atomicAdd(&counters[CNT_RAYS], static_cast<uint32_t>(__popc(tm)));
      rb = __shfl_sync(0xffffffffu, rb, 0);
      if (trace) rays_unsorted[rb + __popc(tm & ((1u << lane) - 1u))] = pt;
    }
  }
  if (lane == 0 && voxels) atomicAdd(&counters[CNT_VOXELS], voxels);
  __syncthreads();
  if (threadIdx.x < kRayLenBins && s_len[threadIdx.x]) atomicAdd(&ray_hist[2 * kRayBins + threadIdx.x], s_len[threadIdx.x]);
}

// level 2, bigger rows: one CTA per row, sorted in shared memory (or, beyond kBigRowSmem points,
// in place in global memory)
__global__ void __launch_bounds__(kBlock)
voxel_rows_big_kernel(const __grid_constant__ VoxelRows vr, const __grid_constant__ RaycastParams p,
                      const DeviceState* __restrict__ st, const float4* __restrict__ pts,
                      uint32_t* __restrict__ counters, float4* __restrict__ rays_unsorted,
                      uint32_t* __restrict__ ray_hist) {
  __shared__ unsigned long long s_pairs[kBigRowSmem];
  __shared__ uint32_t s_len[kRayLenBins];
  __shared__ uint32_t s_job;
  if (threadIdx.x < kRayLenBins) s_len[threadIdx.x] = 0u;
  const GridGeom g = st->geom;
  const int lane = threadIdx.x & 31;
  const bool rc_ok = geom_is_inside(g, static_cast<double>(p.origin[0]), static_cast<double>(p.origin[1]));
  const uint32_t n_jobs = vr.ctl[1];
  uint32_t voxels = 0;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_job = atomicAdd(&vr.ctl[4], 1u);
    __syncthreads();
    const uint32_t job = s_job;
    if (job >= n_jobs) break;
    const uint32_t row = vr.big_list[job];
    const uint32_t base = vr.block_base[row / kRowBlock] + vr.row_excl[row];
    const uint32_t c = vr.row_cursor[row];
    unsigned long long* data;
    uint32_t P = 32;
    while (P < c) P <<= 1;
    if (c <= kBigRowSmem) {
      for (uint32_t e = threadIdx.x; e < P; e += kBlock) s_pairs[e] = e < c ? vr.pairs[base + e] : ~0ull;
      data = s_pairs;
    } else {
      data = vr.pairs + base;   // in place; indices >= c are virtual +inf
    }
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t e = threadIdx.x; e < P; e += kBlock) {
          const uint32_t l = (j == (k >> 1)) ? (e ^ (k - 1)) : (e ^ j);
          if (l > e && l < c) {   // (e < l < c: both real; padding never moves)
            const unsigned long long a = data[e], b = data[l];
            if (b < a) { data[e] = b; data[l] = a; }
          }
        }
        __syncthreads();
      }
    }
    // one thread per sorted element: heads pick their voxel's representative
    for (uint32_t e0 = 0; e0 < c; e0 += kBlock) {
      const uint32_t e = e0 + threadIdx.x;
      bool head = false, trace = false;
      float4 pt = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (e < c) {
        const uint32_t x = static_cast<uint32_t>(data[e] >> 32);
        head = e == 0 || static_cast<uint32_t>(data[e - 1] >> 32) != x;
        if (head) {
          uint32_t lo = e + 1;
          while (lo < c && static_cast<uint32_t>(data[lo] >> 32) == x) ++lo;
          const uint64_t count = lo - e;
          const uint64_t start = static_cast<uint64_t>(base) + e;
          const uint32_t src = static_cast<uint32_t>(
              data[e + (count == 1 ? 0ull : (count * 7ull + start * 13ull) % count)] & 0xffffffffull);
          if (rc_ok) {
            pt = __ldg(&pts[src]);
            trace = ray_of_representative(p, g, pt, ray_hist, s_len);
          }
        }
      }
      const uint32_t hm = __ballot_sync(0xffffffffu, head);
      const uint32_t tm = __ballot_sync(0xffffffffu, trace);
      voxels += (lane == 0) ? __popc(hm) : 0;
      if (tm) {
        uint32_t rb = 0;
        if (lane == 0) rb = atomicAdd(&counters[CNT_RAYS], static_cast<uint32_t>(__popc(tm)));
        rb = __shfl_sync(0xffffffffu, rb, 0);
        if (trace) rays_unsorted[rb + __popc(tm & ((1u << lane) - 1u))] = pt;
      }
    }
  }
  if (lane == 0 && voxels) atomicAdd(&counters[CNT_VOXELS], voxels);
  __syncthreads();
  if (threadIdx.x < kRayLenBins && s_len[threadIdx.x]) atomicAdd(&ray_hist[2 * kRayBins + threadIdx.x], s_len[threadIdx.x]);
}

constexpr int kRayBatch = 8;

// traceRay (raycasting.cpp:46-139): one thread per traced ray, float32 2-D DDA, expression by
// expression as the reference.  What is stored per cell is not the ray height but its offset
// from the sensor height, x = min(t_exit, 1) * dz <= -0 (only downward rays are traced): the
// reference's value is fl(sz + x), float addition of a constant is monotone, so
// min over rays of fl(sz + x_i) == fl(sz + min_i x_i) exactly — and for non-positive floats the
// raw bit pattern grows as the value falls, so the minimum is an atomicMax on the bits, the
// "no ray" state is 0 and no encoding arithmetic is needed in the loop.  The scratch is laid
// out in LOGICAL cell coordinates (column-major, lin = c * nrows + r): the circular-buffer
// wrap is paid once per cell by the resolve kernel instead of once per DDA step.
//  * BUNDLES: the ray list is ordered (length, azimuth); warp task j = rays [32 j, 32 j + 32).
//    Tasks are dealt to the CTAs round-robin, so every CTA (and SM) gets the same mixture of
//    short and long bundles.
//  * NEAR FIELD in shared memory.  Every ray starts in the sensor's cell, so the cells around
//    the sensor see every ray: each CTA keeps a kNear x kNear window in shared memory (a shared
//    load + a rare shared atomic instead of a global load + L2 atomic on the map's hottest
//    words) and flushes it once at the end.
//  * FAR FIELD batched.  The cell sequence does not depend on memory: kRayBatch cells and exit
//    offsets are computed first, their loads issued back to back, then the compares / atomics.
//    A stored value only moves one way, so a stale (L1-cached) read can at worst cause a
//    needless atomic, never a missed one.  The lanes of a bundle often lower the same cell in
//    the same step: they elect one atomic per cell (match.any + redux.max, warp-uniform).
// x <= 0 in value, but a ray that leaves its first cell at t = -0 yields +0: force the sign bit so
// that a visited cell never reads as 0 ("no ray"); fl(sz + -0) == fl(sz + +0) for any sz != -0
__device__ __forceinline__ uint32_t enc_offset(float x) { return __float_as_uint(x) | 0x80000000u; }

template <int kNear, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
raycast_dda_kernel(const __grid_constant__ RaycastParams p, const DeviceState* __restrict__ st,
                   const float4* __restrict__ rays, uint32_t* __restrict__ counters) {
  extern __shared__ uint32_t near_x[];  // [kNear * kNear], column-major in LOGICAL cells
  constexpr int kNearHalf = kNear / 2;
  const uint32_t n_rays = counters[CNT_RAYS];
  if (n_rays == 0) return;                // also: preconditions failed (no ray was listed)
  const GridGeom g = st->geom;
  const int nrows = g.rows, ncols = g.cols;
  const float resolution = static_cast<float>(g.res);
  const float sx = p.origin[0], sy = p.origin[1], sz = p.origin[2];
  const float origin_x = static_cast<float>(g.pos[0]) + nrows * resolution * 0.5f;
  const float origin_y = static_cast<float>(g.pos[1]) + ncols * resolution * 0.5f;
  const float gr0 = (origin_x - sx) / resolution;
  const float gc0 = (origin_y - sy) / resolution;
  const int r_s = static_cast<int>(floorf(gr0));  // the sensor's cell: every ray starts here
  const int c_s = static_cast<int>(floorf(gc0));
  const int win_r0 = r_s - kNearHalf, win_c0 = c_s - kNearHalf;
  const int max_steps = nrows + ncols;
  const int lane = threadIdx.x & 31;
  uint32_t* __restrict__ far_x = p.ray_min_enc;
  // float32 start cell inside the map (the precondition tests the origin in float64; the two
  // can disagree by a rounding at the map's edge — those scans take the generic loop below)
  const bool start_inside = static_cast<unsigned>(r_s) < static_cast<unsigned>(nrows) &&
                            static_cast<unsigned>(c_s) < static_cast<unsigned>(ncols);

  for (int h = threadIdx.x; h < kNear * kNear; h += kBlock) near_x[h] = 0u;
  __syncthreads();

  // SEGMENTS.  A DDA is a serial chain (~850 steps for the longest rays here), and a kernel of
  // one thread per whole ray lasts as long as its longest chain while most SMs idle.  So a ray is
  // cut into segments of kSeg steps, each its own task: the task for segment k first replays the
  // k * kSeg earlier steps with the bare recurrence (compare + add: no memory, no bookkeeping —
  // t_exit and the position are monotone along a ray, so "still running" needs to be checked only
  // at the end of the replay), then walks its kSeg cells in full.  Same arithmetic, same cells,
  // ~1.5x the instructions, but the longest chain drops several-fold.
  const uint32_t n_bundles = (n_rays + 31u) >> 5;
  const int kSeg = p.seg_len;
  const uint32_t k_max = static_cast<uint32_t>((max_steps + kSeg - 1) / kSeg);
  const uint32_t n_tasks = n_bundles * k_max;
  for (;;) {
    uint32_t task = 0;
    if (lane == 0) task = atomicAdd(&counters[CNT_RAY_WORK], 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= n_tasks) break;
    // far segments first: they carry the longest replay
    const int kseg = static_cast<int>(k_max - 1u - task / n_bundles);
    const uint32_t i = (task % n_bundles) * 32u + lane;
    const bool have = i < n_rays;
    const float4 pt = have ? __ldg(&rays[i]) : make_float4(sx, sy, sz, 0.0f);
    const float dz = pt.z - sz;
    const float gr1 = (origin_x - pt.x) / resolution;
    const float gc1 = (origin_y - pt.y) / resolution;
    const float dr = gr1 - gr0;
    const float dc = gc1 - gc0;
    // cells a ray that stays in the map visits: |rows crossed| + |columns crossed| + 1 (a couple
    // more or fewer at float edges: the margin below only decides whether the task is looked at)
    const int n_est = have ? static_cast<int>(fabsf(floorf(gr1) - static_cast<float>(r_s)) +
                                              fabsf(floorf(gc1) - static_cast<float>(c_s))) + 4 : 0;
    const int s0 = kseg * kSeg;
    if (!__any_sync(0xffffffffu, n_est > s0)) continue;   // the whole bundle ends before this segment
    int r = r_s, c = c_s;
    int step_r, step_c;
    float t_r, t_c, td_r, td_c;  // t_max_r / t_max_c / t_delta_r / t_delta_c
    if (fabsf(dr) > 1e-8f) {
      step_r = (dr > 0) ? 1 : -1;
      const float boundary = (step_r > 0) ? (r + 1.0f) : static_cast<float>(r);
      t_r = (boundary - gr0) / dr;
      td_r = static_cast<float>(step_r) / dr;
    } else {
      step_r = 0;
      t_r = 1e30f;
      td_r = 1e30f;
    }
    if (fabsf(dc) > 1e-8f) {
      step_c = (dc > 0) ? 1 : -1;
      const float boundary = (step_c > 0) ? (c + 1.0f) : static_cast<float>(c);
      t_c = (boundary - gc0) / dc;
      td_c = static_cast<float>(step_c) / dc;
    } else {
      step_c = 0;
      t_c = 1e30f;
      td_c = 1e30f;
    }

    if (!start_inside) {
      // generic traversal (rare: float32 start cell outside the map), whole ray in one task, one
      // cell per round trip, every test of the reference's loop.  The map is convex: once a ray
      // has left it, it never comes back (the reference keeps stepping through cells that fail
      // its bounds test; stopping there changes nothing).
      if (kseg != 0) continue;
      bool alive = have, was_inside = false;
      for (int s = 0; alive && s < max_steps; ++s) {
        const bool row = t_r < t_c;
        const float t_exit = row ? t_r : t_c;   // == min(t_max_r, t_max_c)
        if (static_cast<unsigned>(r) < static_cast<unsigned>(nrows) &&
            static_cast<unsigned>(c) < static_cast<unsigned>(ncols)) {
          was_inside = true;
          atomicMax(&far_x[c * nrows + r], enc_offset(fminf(t_exit, 1.0f) * dz));
        } else if (was_inside) {
          break;
        }
        if (t_exit >= 1.0f) break;
        if (row) { r += step_r; t_r += td_r; } else { c += step_c; t_c += td_c; }
      }
      continue;
    }

    // ── replay of the s0 earlier steps: the bare recurrence ──
    float t_prev = 0.0f;   // t_exit of the last replayed step
    int nr = 0;            // row steps among them
    for (int s = 0; s < s0; ++s) {
      const bool row = t_r < t_c;
      t_prev = row ? t_r : t_c;
      if (row) { t_r += td_r; ++nr; } else { t_c += td_c; }
    }
    r = r_s + step_r * nr;
    c = c_s + step_c * (s0 - nr);
    // Start cell inside the map and monotone movement: every cell reached before the ray steps
    // onto row r_out or column c_out is inside, so the bounds test is two compares per step and
    // the reference's step cap (nrows + ncols) can never bind first.
    const int r_out = step_r > 0 ? nrows : -1;      // step_r == 0: r never changes, never equal
    const int c_out = step_c > 0 ? ncols : -1;
    const int dlin_c = step_c * nrows;
    bool alive = have && t_prev < 1.0f && static_cast<unsigned>(r) < static_cast<unsigned>(nrows) &&
                 static_cast<unsigned>(c) < static_cast<unsigned>(ncols);
    int lin = c * nrows + r;
    int left = kSeg;       // steps of this segment still to walk

    // ── near field: shared-memory window around the sensor (first segment only).  The lanes of a
    // bundle sit on the same window cell most of the time: when they do, the warp writes its best
    // value with ONE shared atomic instead of 32 conflicting ones ──
    if (kseg == 0) {
      for (;;) {
        const int wr = r - win_r0, wc = c - win_c0;
        const bool inwin = alive && left > 0 && static_cast<unsigned>(wr) < static_cast<unsigned>(kNear) &&
                           static_cast<unsigned>(wc) < static_cast<unsigned>(kNear);
        if (!__any_sync(0xffffffffu, inwin)) break;   // every lane has left the window (or ended)
        const bool row = t_r < t_c;
        const float t_exit = row ? t_r : t_c;
        const uint32_t e = enc_offset(fminf(t_exit, 1.0f) * dz);
        const int w = inwin ? wc * kNear + wr : 0;
        const bool want = inwin && e > near_x[w];
        const uint32_t wm = __ballot_sync(0xffffffffu, want);
        if (wm) {
          const int w0 = __shfl_sync(0xffffffffu, w, __ffs(wm) - 1);
          if (__all_sync(0xffffffffu, !want || w == w0)) {
            const uint32_t best = __reduce_max_sync(0xffffffffu, want ? e : 0u);
            if (lane == __ffs(wm) - 1) atomicMax(&near_x[w0], best);
          } else if (want) {
            atomicMax(&near_x[w], e);
          }
        }
        if (inwin) {
          --left;
          if (t_exit >= 1.0f) {
            alive = false;
          } else {
            if (row) { r += step_r; t_r += td_r; lin += step_r; } else { c += step_c; t_c += td_c; lin += dlin_c; }
            alive = r != r_out && c != c_out;
          }
        }
      }
    }

    // ── far field: global scratch, kRayBatch cells per memory round trip ──
    alive = alive && left > 0;
    while (__any_sync(0xffffffffu, alive)) {
      int idx[kRayBatch];
      uint32_t ev[kRayBatch];
      uint32_t cv[kRayBatch];
#pragma unroll
      for (int k = 0; k < kRayBatch; ++k) {
        const bool row = t_r < t_c;
        const float t_exit = row ? t_r : t_c;
        ev[k] = enc_offset(fminf(t_exit, 1.0f) * dz);
        idx[k] = alive ? lin : -1;
        r += row ? step_r : 0;
        c += row ? 0 : step_c;
        lin += row ? step_r : dlin_c;
        t_r = row ? t_r + td_r : t_r;
        t_c = row ? t_c : t_c + td_c;
        --left;
        // the ray ended in that cell (t >= 1), has just stepped out of the map, or the segment is done
        alive = alive && t_exit < 1.0f && r != r_out && c != c_out && left > 0;
      }
      if (p.tune_ld_cg) {
#pragma unroll
        for (int k = 0; k < kRayBatch; ++k) cv[k] = idx[k] >= 0 ? __ldcg(&far_x[idx[k]]) : 0xffffffffu;
      } else {
#pragma unroll
        for (int k = 0; k < kRayBatch; ++k) cv[k] = idx[k] >= 0 ? far_x[idx[k]] : 0xffffffffu;
      }
#pragma unroll
      for (int k = 0; k < kRayBatch; ++k) {
        const bool want = ev[k] > cv[k];
        if (p.tune_elect == 0) {
          if (want) atomicMax(&far_x[idx[k]], ev[k]);
        } else {
          // bundle fast path: when every writing lane is on the same cell, one lane writes the warp's best
          const uint32_t wm = __ballot_sync(0xffffffffu, want);
          if (wm) {
            const int i0 = __shfl_sync(0xffffffffu, idx[k], __ffs(wm) - 1);
            if (__all_sync(0xffffffffu, !want || idx[k] == i0)) {
              const uint32_t best = __reduce_max_sync(0xffffffffu, want ? ev[k] : 0u);
              if (lane == __ffs(wm) - 1) atomicMax(&far_x[i0], best);
            } else if (want) {
              atomicMax(&far_x[idx[k]], ev[k]);
            }
          }
        }
      }
    }
  }

  // ── flush the window: every near-field cell this CTA lowered, once ──
  __syncthreads();
  for (int h = threadIdx.x; h < kNear * kNear; h += kBlock) {
    const uint32_t v = near_x[h];
    if (v == 0u) continue;
    const int r = win_r0 + (h % kNear), c = win_c0 + (h / kNear);  // in the map: only such cells are written
    uint32_t* slot = &far_x[c * nrows + r];
    if (v > __ldcg(slot)) atomicMax(slot, v);
  }
}

// map.clear(raycasting) (:242) + the observed log-odds updates + resolveGhostCells
// (:188-214), one thread per cell; also rearms the scratch for the next scan.  `hits` is
// indexed like the layers (buffer cell), the ray minima by LOGICAL cell (see the DDA kernel).
__global__ void __launch_bounds__(kBlock)
raycast_resolve_kernel(const __grid_constant__ RaycastParams p, const DeviceState* __restrict__ st,
                       const __grid_constant__ LayerTable lt,
                       const uint32_t* __restrict__ counters, size_t n_cells) {
  if (counters[CNT_RC_SKIP]) return;  // applyRaycasting returned before touching anything
  const int nrows = st->geom.rows, ncols = st->geom.cols;
  const int start_r = st->geom.start[0], start_c = st->geom.start[1];
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n_cells; i += stride) {
    const int bc = static_cast<int>(i / static_cast<size_t>(nrows));
    const int br = static_cast<int>(i - static_cast<size_t>(bc) * nrows);
    int lr = br - start_r;  // buffer = (logical + start) % size
    if (lr < 0) lr += nrows;
    int lc = bc - start_c;
    if (lc < 0) lc += ncols;
    uint32_t* enc_slot = &p.ray_min_enc[static_cast<size_t>(lc) * nrows + lr];
    const uint32_t hits = p.hits[i];
    const uint32_t enc = *enc_slot;
    if (hits == 0 && enc == 0u) {
      p.raycasting[i] = nanf_();
      continue;
    }
    p.hits[i] = 0;
    *enc_slot = 0u;
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
    if (enc != 0u) {
      rmin = p.origin[2] + __uint_as_float(enc);  // start.z() + min(t_exit, 1) * dz of the lowest ray
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

// The same sweep on one ROW STRIPE of a GLOBAL map (start index 0): the logical rows just above
// and below the stripe belong to the neighbouring ranks and arrive as two rows of `cols` floats
// (null at the map's border) — the halo the sharded driver exchanges between sweeps.
__global__ void __launch_bounds__(kBlock)
inpaint_stripe_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows_local, int cols,
                      const float* __restrict__ above, const float* __restrict__ below, int min_valid) {
  const size_t n = static_cast<size_t>(rows_local) * cols;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const float v = src[i];
    float out = v;
    if (isnan(v)) {
      const int c = static_cast<int>(i / rows_local);
      const int r = static_cast<int>(i - static_cast<size_t>(c) * rows_local);
      float sum = 0.0f;
      int count = 0;
      for (int dr = -1; dr <= 1; ++dr) {
        for (int dc = -1; dc <= 1; ++dc) {
          if (dr == 0 && dc == 0) continue;
          const int nr = r + dr, nc = c + dc;
          if (nc < 0 || nc >= cols) continue;
          float val;
          if (nr < 0) {
            if (!above) continue;
            val = above[nc];
          } else if (nr >= rows_local) {
            if (!below) continue;
            val = below[nc];
          } else {
            val = src[static_cast<size_t>(nc) * rows_local + nr];
          }
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

// ── circular neighbourhoods: nanogrid region(radius) / neighbors() are not in the reference
// tree (un-vendored nanoGrid), so they are restated from the call sites:
// offsets with (float)((dr^2+dc^2) res^2) <= radius^2, centre included, dr outer / dc inner,
// restricted to in-bounds LOGICAL cells ──
constexpr int kMaxRegionHalf = 8;
constexpr int kMaxRegionCells = (2 * kMaxRegionHalf + 1) * (2 * kMaxRegionHalf + 1);  // 289
constexpr int kSelectMax = 16;

// weighted quantile of samples already sorted by value (SimpleWeightedECDF::quantile,
// uncertainty_fusion.cpp:62-90)
__device__ __forceinline__ float weighted_quantile(const float* v, const float* w, int n, float p) {
  if (n == 0) return nanf_();
  if (n == 1) return v[0];
  float total = 0.0f;
  for (int k = 0; k < n; ++k) total += w[k];
  if (total <= 0.0f) return nanf_();
  const float target = p * total;
  float cumulative = 0.0f;
  for (int k = 0; k < n; ++k) {
    cumulative += w[k];
    if (cumulative >= target) return v[k];
  }
  return v[n - 1];
}

// stable insertion of (val, wt) into arrays sorted by val
__device__ __forceinline__ void sorted_insert(float* v, float* w, int n, float val, float wt) {
  int j = n;
  while (j > 0 && val < v[j - 1]) {
    v[j] = v[j - 1];
    w[j] = w[j - 1];
    --j;
  }
  v[j] = val;
  w[j] = wt;
}

// applyUncertaintyFusion (fastdem/src/uncertainty_fusion.cpp:103-186): bilateral weights
// (Gaussian in distance x inverse bound range) -> weighted quantiles of the neighbours'
// lower / upper bounds.  Reads the snapshots, writes the layers (double buffer).
__global__ void __launch_bounds__(128)
uncertainty_fusion_kernel(const float* __restrict__ upper_in, const float* __restrict__ lower_in,
                          float* __restrict__ upper_out, float* __restrict__ lower_out,
                          const DeviceState* __restrict__ st, int half, float radius_sq,
                          float inv_2sigma_sq, float q_lower, float q_upper, int min_valid,
                          int rows_local, int cols) {
  const GridGeom g = st->geom;
  const size_t n = static_cast<size_t>(rows_local) * cols;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float cu = upper_in[i], cl = lower_in[i];
  if (!isfinite(cu) || !isfinite(cl)) return;
  const int bc = static_cast<int>(i / rows_local);
  const int br = static_cast<int>(i - static_cast<size_t>(bc) * rows_local);
  const int lr = wrap_index(br - g.start[0] + g.rows, g.rows);
  const int lc = wrap_index(bc - g.start[1] + g.cols, g.cols);
  float lo_v[kMaxRegionCells], lo_w[kMaxRegionCells], up_v[kMaxRegionCells], up_w[kMaxRegionCells];
  int ns = 0, valid = 0;
  const double res2 = g.res * g.res;
  for (int dr = -half; dr <= half; ++dr) {
    for (int dc = -half; dc <= half; ++dc) {
      const float dist_sq = static_cast<float>(static_cast<double>(dr * dr + dc * dc) * res2);
      if (!(dist_sq <= radius_sq)) continue;
      const int nr = lr + dr, nc = lc + dc;
      if (nr < 0 || nr >= g.rows || nc < 0 || nc >= g.cols) continue;
      const size_t nl = static_cast<size_t>(wrap_index(nc + g.start[1], g.cols)) * rows_local +
                        wrap_index(nr + g.start[0], g.rows);
      const float nu = upper_in[nl], nlo = lower_in[nl];
      if (!isfinite(nu) || !isfinite(nlo)) continue;
      const float w_spatial = expf(-dist_sq * inv_2sigma_sq);
      const float range = nu - nlo;
      const float w_range = 1.0f / (range + 1e-4f);
      const float w = w_spatial * w_range;
      if (w > 1e-6f) {
        sorted_insert(lo_v, lo_w, ns, nlo, w);
        sorted_insert(up_v, up_w, ns, nu, w);
        ++ns;
      }
      ++valid;
    }
  }
  if (valid >= min_valid) {
    const float lower = weighted_quantile(lo_v, lo_w, ns, q_lower);
    const float upper = weighted_quantile(up_v, up_w, ns, q_upper);
    if (isfinite(lower) && isfinite(upper)) {
      upper_out[i] = upper;
      lower_out[i] = lower;
    }
  }
}

// ── 3x3 symmetric eigen-decomposition, the closed form Eigen's
// SelfAdjointEigenSolver<Matrix3f>::computeDirect uses (nanopcl::geometry::computePCA,
// nanopcl/geometry/impl/pca.hpp:67-90): shift by trace/3, scale to [-1,1], trigonometric
// roots, eigenvectors from row cross products. ──
struct Eig3 {
  float val[3];
  float vec[3][3];
};
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float sqnorm3v(const float* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
__device__ __forceinline__ void eig3_kernel(const float m[3][3], float* res, float* representative) {
  int i0 = 0;
  float best = fabsf(m[0][0]);
#pragma unroll
  for (int i = 1; i < 3; ++i)
    if (fabsf(m[i][i]) > best) { best = fabsf(m[i][i]); i0 = i; }
  float c0[3], c1[3], a[3], b[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    representative[i] = (i0 == 0) ? m[i][0] : (i0 == 1) ? m[i][1] : m[i][2];
    a[i] = (i0 == 0) ? m[i][1] : (i0 == 1) ? m[i][2] : m[i][0];  // column (i0+1)%3
    b[i] = (i0 == 0) ? m[i][2] : (i0 == 1) ? m[i][0] : m[i][1];  // column (i0+2)%3
  }
  cross3(representative, a, c0);
  cross3(representative, b, c1);
  const float n0 = sqnorm3v(c0), n1 = sqnorm3v(c1);
  if (n0 > n1) {
    const float s = sqrtf(n0);
#pragma unroll
    for (int i = 0; i < 3; ++i) res[i] = c0[i] / s;
  } else {
    const float s = sqrtf(n1);
#pragma unroll
    for (int i = 0; i < 3; ++i) res[i] = c1[i] / s;
  }
}
__device__ __forceinline__ Eig3 eig3_direct(const float cov[3][3]) {
  Eig3 out;
  constexpr float eps = 1.1920929e-07f;
  const float shift = (cov[0][0] + cov[1][1] + cov[2][2]) / 3.0f;
  float m[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) m[i][j] = (i >= j) ? cov[i][j] : cov[j][i];
#pragma unroll
  for (int i = 0; i < 3; ++i) m[i][i] -= shift;
  float scale = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) scale = fmaxf(scale, fabsf(m[i][j]));
  if (scale > 0.0f) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) m[i][j] /= scale;
  }
  const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[1][0] * m[2][0] * m[2][1] -
                   m[0][0] * m[2][1] * m[2][1] - m[1][1] * m[2][0] * m[2][0] - m[2][2] * m[1][0] * m[1][0];
  const float c1 = m[0][0] * m[1][1] - m[1][0] * m[1][0] + m[0][0] * m[2][2] - m[2][0] * m[2][0] +
                   m[1][1] * m[2][2] - m[2][1] * m[2][1];
  const float c2 = m[0][0] + m[1][1] + m[2][2];
  const float inv3 = 1.0f / 3.0f, sqrt3 = sqrtf(3.0f);
  const float c2_3 = c2 * inv3;
  float a_3 = (c2 * c2_3 - c1) * inv3;
  a_3 = fmaxf(a_3, 0.0f);
  const float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
  float q = a_3 * a_3 * a_3 - half_b * half_b;
  q = fmaxf(q, 0.0f);
  const float rho = sqrtf(a_3);
  const float theta = atan2f(sqrtf(q), half_b) * inv3;
  const float ct = cosf(theta), sn = sinf(theta);
  float ev[3] = {c2_3 - rho * (ct + sqrt3 * sn), c2_3 - rho * (ct - sqrt3 * sn), c2_3 + 2.0f * rho * ct};
  float V0[3] = {1, 0, 0}, V1[3] = {0, 1, 0}, V2[3] = {0, 0, 1};
  if (!((ev[2] - ev[0]) <= eps)) {
    float d0 = ev[2] - ev[1];
    const float d1 = ev[1] - ev[0];
    const bool swapped = d0 > d1;  // Eigen: swap(k, l); d0 = d1;
    if (swapped) d0 = d1;
    float* Vk = swapped ? V2 : V0;
    float* Vl = swapped ? V0 : V2;
    const float ek = swapped ? ev[2] : ev[0];
    const float el = swapped ? ev[0] : ev[2];
    float tmp[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) tmp[i][j] = m[i][j];
#pragma unroll
    for (int i = 0; i < 3; ++i) tmp[i][i] -= ek;
    eig3_kernel(tmp, Vk, Vl);
    if (d0 <= 2.0f * eps * d1) {
      const float dot = Vk[0] * Vl[0] + Vk[1] * Vl[1] + Vk[2] * Vl[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) Vl[i] -= dot * Vl[i];
      const float nrm = sqrtf(sqnorm3v(Vl));
#pragma unroll
      for (int i = 0; i < 3; ++i) Vl[i] /= nrm;
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) tmp[i][j] = m[i][j];
#pragma unroll
      for (int i = 0; i < 3; ++i) tmp[i][i] -= el;
      float dummy[3];
      eig3_kernel(tmp, Vl, dummy);
    }
    float c[3];
    cross3(V2, V0, c);
    const float nrm = sqrtf(sqnorm3v(c));
#pragma unroll
    for (int i = 0; i < 3; ++i) V1[i] = c[i] / nrm;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) out.val[k] = ev[k] * scale + shift;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    out.vec[0][i] = V0[i];
    out.vec[1][i] = V1[i];
    out.vec[2][i] = V2[i];
  }
  return out;
}

struct FeatureLayers {
  float *step, *slope, *roughness, *curvature, *nx, *ny, *nz;
};

// applyFeatureExtraction (fastdem/src/feature_extraction.cpp:28-118): local PCA of the
// elevation patch -> normal, slope, roughness, curvature; percentile range -> step
__global__ void __launch_bounds__(128)
feature_extraction_kernel(const float* __restrict__ elev, const FeatureLayers out,
                          const DeviceState* __restrict__ st, int half, float radius_sq,
                          int min_valid, float p_lo, float p_hi, int rows_local, int cols,
                          int keep_small, int keep_large) {
  const GridGeom g = st->geom;
  const size_t n = static_cast<size_t>(rows_local) * cols;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float center_z = elev[i];
  if (!isfinite(center_z)) return;
  const int bc = static_cast<int>(i / rows_local);
  const int br = static_cast<int>(i - static_cast<size_t>(bc) * rows_local);
  const int lr = wrap_index(br - g.start[0] + g.rows, g.rows);
  const int lc = wrap_index(bc - g.start[1] + g.cols, g.cols);
  const float resf = static_cast<float>(g.res);
  const double res2 = g.res * g.res;
  float sum[3] = {0.0f, 0.0f, 0.0f};
  float sq[3][3] = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
  // The step feature needs two order statistics of the patch's heights.  When both ranks are
  // near the ends (the default 5 % / 95 %), only the keep_small smallest and keep_large largest
  // values are tracked (the host sizes them from the region's cell count); otherwise the whole
  // patch is kept sorted.
  const bool select = keep_small > 0;
  float z_vals[kMaxRegionCells];
  float z_small[kSelectMax], z_large[kSelectMax];
  int n_small = 0, n_large = 0;
  int count = 0;
  for (int dr = -half; dr <= half; ++dr) {
    for (int dc = -half; dc <= half; ++dc) {
      const float dist_sq = static_cast<float>(static_cast<double>(dr * dr + dc * dc) * res2);
      if (!(dist_sq <= radius_sq)) continue;
      const int nr = lr + dr, nc = lc + dc;
      if (nr < 0 || nr >= g.rows || nc < 0 || nc >= g.cols) continue;
      const float nz = elev[static_cast<size_t>(wrap_index(nc + g.start[1], g.cols)) * rows_local +
                            wrap_index(nr + g.start[0], g.rows)];
      if (!isfinite(nz)) continue;
      // world-frame displacement (grid_map: row -> -x, col -> -y)
      const float d[3] = {-dr * resf, -dc * resf, nz - center_z};
#pragma unroll
      for (int a = 0; a < 3; ++a) sum[a] += d[a];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) sq[a][b] += d[a] * d[b];
      if (select) {
        if (n_small < keep_small || nz < z_small[n_small - 1]) {  // ascending, smallest first
          int j = n_small < keep_small ? n_small++ : n_small - 1;
          while (j > 0 && nz < z_small[j - 1]) {
            z_small[j] = z_small[j - 1];
            --j;
          }
          z_small[j] = nz;
        }
        if (n_large < keep_large || nz > z_large[n_large - 1]) {  // descending, largest first
          int j = n_large < keep_large ? n_large++ : n_large - 1;
          while (j > 0 && nz > z_large[j - 1]) {
            z_large[j] = z_large[j - 1];
            --j;
          }
          z_large[j] = nz;
        }
      } else {
        // keep z_vals sorted (the reference sorts afterwards; same multiset)
        int j = count;
        while (j > 0 && nz < z_vals[j - 1]) {
          z_vals[j] = z_vals[j - 1];
          --j;
        }
        z_vals[j] = nz;
      }
      ++count;
    }
  }
  if (count < min_valid) return;
  const float inv_n = 1.0f / static_cast<float>(count);
  float mean[3], cov[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) mean[a] = sum[a] * inv_n;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) cov[a][b] = sq[a][b] * inv_n - mean[a] * mean[b];
  const float trace = cov[0][0] + cov[1][1] + cov[2][2];
  if (trace < 1.1920929e-07f) return;  // computePCA: degenerate covariance -> invalid
  const Eig3 pca = eig3_direct(cov);
  if (pca.val[1] < 1e-8f) return;      // collinear patch
  float normal[3] = {pca.vec[0][0], pca.vec[0][1], pca.vec[0][2]};
  if (normal[2] < 0.0f) {
    normal[0] = -normal[0];
    normal[1] = -normal[1];
    normal[2] = -normal[2];
  }
  const int lo = static_cast<int>(p_lo * (count - 1));
  const int hi = static_cast<int>(p_hi * (count - 1));
  // ranks from the top: (count-1) - hi never exceeds the host's bound for the full region
  out.step[i] = select ? z_large[(count - 1) - hi] - z_small[lo] : z_vals[hi] - z_vals[lo];
  out.slope[i] = acosf(fabsf(normal[2])) * 180.0f / 3.14159265358979323846f;
  out.roughness[i] = sqrtf(pca.val[0]);
  out.curvature[i] = (trace > 0.0f) ? fabsf(pca.val[0] / trace) : 0.0f;
  out.nx[i] = normal[0];
  out.ny[i] = normal[1];
  out.nz[i] = normal[2];
}

// ── map -> sensor_msgs/PointCloud2 body (toPointCloud2Impl,
// fastdem/include/fastdem/bridge/ros/impl.hpp:29-174): one point per cell with a finite
// elevation; columns outer, rows inner, both starting at sub_start and wrapping around the
// circular buffer.  One warp per sub-region column: count -> scan -> ordered write. ──
struct PackParams {
  const float* elev;
  const float* field[kMaxLayers];  // visible float layers, then (optionally) the colour bits
  int32_t n_fields;                // entries in field[]
  int32_t rows, cols;              // buffer size (unsharded map)
  int32_t sub_r0, sub_c0, sub_rows, sub_cols;
};

__global__ void __launch_bounds__(kBlock)
pack_count_kernel(const __grid_constant__ PackParams p, uint32_t* __restrict__ col_count) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.sub_cols) return;
  const size_t base = static_cast<size_t>((p.sub_c0 + warp) % p.cols) * p.rows;
  uint32_t n = 0;
  for (int i = lane; i < p.sub_rows; i += 32)
    n += isfinite(p.elev[base + (p.sub_r0 + i) % p.rows]) ? 1u : 0u;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) n += __shfl_down_sync(0xffffffffu, n, d);
  if (lane == 0) col_count[warp] = n;
}

// exclusive scan of the per-column counts (<= a few thousand entries): one CTA
__global__ void __launch_bounds__(1024)
pack_scan_kernel(const uint32_t* __restrict__ col_count, uint32_t* __restrict__ col_offset, int n,
                 uint32_t* __restrict__ total) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int j = base + threadIdx.x;
    const uint32_t v = j < n ? col_count[j] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wprefix = 0, all = 0;
    for (int w = 0; w < 32; ++w) {
      const uint32_t x = s_warp[w];
      if (w < warp) wprefix += x;
      all += x;
    }
    if (j < n) col_offset[j] = s_running + wprefix + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) s_running += all;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_running;
}

__global__ void __launch_bounds__(kBlock)
pack_write_kernel(const __grid_constant__ PackParams p, const DeviceState* __restrict__ st,
                  const uint32_t* __restrict__ col_offset, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.sub_cols) return;
  const GridGeom g = st->geom;
  const double origin_x = g.pos[0] + g.len[0] / 2.0 - g.res / 2.0;
  const double origin_y = g.pos[1] + g.len[1] / 2.0 - g.res / 2.0;
  const int bc = (p.sub_c0 + warp) % p.cols;
  const int uc = (bc - g.start[1] + p.cols) % p.cols;
  const float y = static_cast<float>(origin_y - uc * g.res);
  const size_t base = static_cast<size_t>(bc) * p.rows;
  const int step = 3 + p.n_fields;  // floats per point
  uint32_t pos = col_offset[warp];
  for (int i0 = 0; i0 < p.sub_rows; i0 += 32) {
    const int i = i0 + lane;
    float z = nanf_();
    size_t idx = 0;
    int br = 0;
    if (i < p.sub_rows) {
      br = (p.sub_r0 + i) % p.rows;
      idx = base + br;
      z = p.elev[idx];
    }
    const bool ok = isfinite(z);
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int ur = (br - g.start[0] + p.rows) % p.rows;
      float* o = out + static_cast<size_t>(pos + __popc(m & ((1u << lane) - 1u))) * step;
      o[0] = static_cast<float>(origin_x - ur * g.res);
      o[1] = y;
      o[2] = z;
      for (int f = 0; f < p.n_fields; ++f) o[3 + f] = p.field[f][idx];
    }
    pos += __popc(m);
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
void launch_voxel_keys32(const float4* pm, uint32_t n, float inv_voxel, const VoxelBox& box,
                         uint32_t* keys, uint32_t* vals, uint32_t* counters, cudaStream_t s,
                         LaunchCounter& lc) {
  if (n == 0) return;
  voxel_keys32_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(pm, n, inv_voxel, box, keys, vals,
                                                                   counters);
  ++lc.mine;
}
void launch_voxel_select(const uint64_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                         uint32_t* counters, uint32_t* out_sel, cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  voxel_select_kernel<uint64_t><<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(
      sorted_keys, sorted_vals, n, kInvalidVoxel, counters, out_sel);
  ++lc.mine;
}
void launch_voxel_select32(const uint32_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                           uint32_t invalid_key, uint32_t* counters, uint32_t* out_sel,
                           cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  voxel_select_kernel<uint32_t><<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(
      sorted_keys, sorted_vals, n, invalid_key, counters, out_sel);
  ++lc.mine;
}
// histogram [kRayBins] | cursors [kRayBins] | rays per length bin [kRayLenBins]
size_t ray_sort_scratch_words() { return 2 * static_cast<size_t>(kRayBins) + kRayLenBins; }
static_assert(kRayLenBins == 32, "ray_bin_scan_kernel sums the length-bin totals with one warp");

static void launch_ray_bundle_sort(const RaySortScratch& rs, uint32_t n_max, const uint32_t* counters,
                                   cudaStream_t s, LaunchCounter& lc) {
  ray_bin_scan_kernel<<<kRayLenBins, kRayAzBins, 0, s>>>(rs.hist, rs.hist + kRayBins);
  const uint32_t want = (n_max + kBlock - 1) / kBlock;
  ray_bin_scatter_kernel<<<want < 148u * 4u ? want : 148u * 4u, kBlock, 0, s>>>(rs.unsorted, rs.rays,
                                                                              rs.hist + kRayBins, counters);
  lc.mine += 2;
}
void launch_voxel_select_rays32(const uint32_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                                uint32_t invalid_key, const RaycastParams& p, const DeviceState* st,
                                const float4* pts, uint32_t* counters, const RaySortScratch& rs,
                                cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  voxel_select_rays_kernel<uint32_t><<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(
      sorted_keys, sorted_vals, n, invalid_key, p, st, pts, counters, rs.unsorted, rs.hist);
  ++lc.mine;
  launch_ray_bundle_sort(rs, n, counters, s, lc);
}
void launch_voxel_select_rays64(const uint64_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n,
                                const RaycastParams& p, const DeviceState* st, const float4* pts,
                                uint32_t* counters, const RaySortScratch& rs, cudaStream_t s,
                                LaunchCounter& lc) {
  if (n == 0) return;
  voxel_select_rays_kernel<uint64_t><<<(n + kBlock - 1) / kBlock, kBlock, 0, s>>>(
      sorted_keys, sorted_vals, n, kInvalidVoxel, p, st, pts, counters, rs.unsorted, rs.hist);
  ++lc.mine;
  launch_ray_bundle_sort(rs, n, counters, s, lc);
}
// voxelGrid(ANY) + processScan's per-point part through the MSD sort (no library launches)
void launch_voxel_select_rays_msd(const float4* pm, uint32_t n, float inv_voxel, const VoxelBox& box,
                                  const VoxelRowsScratch& vs, const RaycastParams& p, const DeviceState* st,
                                  uint32_t* counters, const RaySortScratch& rs, cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  VoxelRows vr{};
  vr.bx = static_cast<uint32_t>(box.bx);
  vr.n_rows = 1u << (box.by + box.bz);
  vr.n_blocks = (vr.n_rows + kRowBlock - 1) / kRowBlock;
  vr.invalid_key = box.invalid_key;
  uint32_t* w = vs.words;
  vr.row_count = w;  w += vs.rows_cap;
  vr.row_excl = w;   w += vs.rows_cap;
  vr.row_cursor = w; w += vs.rows_cap;
  vr.small_list = w; w += vs.rows_cap;
  vr.big_list = w;   w += vs.rows_cap;
  vr.block_sum = w;  w += vs.rows_cap / kRowBlock + 1;
  vr.block_base = w; w += vs.rows_cap / kRowBlock + 1;
  vr.ctl = w;
  vr.vkey = vs.vkey;
  vr.pairs = vs.pairs;
  const uint32_t grid_n = (n + kBlock - 1) / kBlock;
  voxel_rows_kernel<<<grid_n, kBlock, 0, s>>>(pm, n, inv_voxel, box, vr, counters);
  voxel_row_scan_kernel<<<vr.n_blocks, kBlock, 0, s>>>(vr);
  voxel_scatter_kernel<<<grid_n, kBlock, 0, s>>>(n, vr);
  voxel_rows_big_kernel<<<148 * 4, kBlock, 0, s>>>(vr, p, st, pm, counters, rs.unsorted, rs.hist);
  voxel_rows_small_kernel<<<148 * 8, kBlock, 0, s>>>(vr, p, st, pm, counters, rs.unsorted, rs.hist);
  lc.mine += 5;
  launch_ray_bundle_sort(rs, n, counters, s, lc);
}
size_t voxel_rows_scratch_words(uint32_t rows_cap) {
  return 5 * static_cast<size_t>(rows_cap) + 2 * (static_cast<size_t>(rows_cap) / kRowBlock + 1) + 8;
}
void launch_raycast_dda(const RaycastParams& p, const DeviceState* st, const float4* rays,
                        uint32_t n_max, uint32_t* counters, cudaStream_t s, LaunchCounter& lc) {
  if (n_max == 0) return;
  // near-field window size / persistent CTAs per SM (FDEM_RAY_NEAR=64|96|128 overrides)
  static int near = -1;
  if (near < 0) {
    const char* e = std::getenv("FDEM_RAY_NEAR");
    near = e ? std::atoi(e) : 64;
    if (near != 64 && near != 96 && near != 128) near = 64;
  }
  const uint32_t chunks = (n_max + 31) / 32;  // warps fetch 32 rays at a time
  auto launch = [&](auto kernel, int kn, uint32_t ctas_per_sm) {
    const size_t smem = sizeof(uint32_t) * kn * kn;
    // opt-in to > 48 KiB of dynamic shared memory (a per-device attribute: set once per device)
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    static const int ctas_env = [] { const char* e = std::getenv("FDEM_RAY_CTAS"); return e ? std::atoi(e) : 0; }();
    if (ctas_env > 0 && static_cast<uint32_t>(ctas_env) < ctas_per_sm) ctas_per_sm = static_cast<uint32_t>(ctas_env);
    const uint32_t want = (chunks + (kBlock / 32) - 1) / (kBlock / 32) * 4u;   // ~4 segment tasks per bundle
    const uint32_t grid = want < 148u * ctas_per_sm ? want : 148u * ctas_per_sm;
    kernel<<<grid, kBlock, smem, s>>>(p, st, rays, counters);
  };
  static const int occ = [] { const char* e = std::getenv("FDEM_RAY_OCC"); return e ? std::atoi(e) : 0; }();
  if (near == 64 && occ == 8) launch(raycast_dda_kernel<64, 8>, 64, 8u);
  else if (near == 64) launch(raycast_dda_kernel<64, 6>, 64, 6u);
  else if (near == 96) launch(raycast_dda_kernel<96, 4>, 96, 5u);
  else launch(raycast_dda_kernel<128, 3>, 128, 3u);
  ++lc.mine;
}
void launch_raycast_resolve(const RaycastParams& p, const DeviceState* st,
                            const LayerTable& lt, const uint32_t* counters, size_t n_cells,
                            cudaStream_t s, LaunchCounter& lc) {
  raycast_resolve_kernel<<<grid_for(n_cells, kBlock), kBlock, 0, s>>>(p, st, lt, counters, n_cells);
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
int launch_uncertainty_fusion(const float* upper_in, const float* lower_in, float* upper_out,
                              float* lower_out, const DeviceState* st, float radius, double res,
                              float spatial_sigma, float q_lower, float q_upper, int min_valid,
                              cudaStream_t s, LaunchCounter& lc, int rows_local, int cols) {
  const int half = static_cast<int>(ceil(static_cast<double>(radius) / res));
  if (half > kMaxRegionHalf) return 1;
  const size_t n = static_cast<size_t>(rows_local) * cols;
  if (n == 0) return 0;
  uncertainty_fusion_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, s>>>(
      upper_in, lower_in, upper_out, lower_out, st, half, radius * radius,
      1.0f / (2.0f * spatial_sigma * spatial_sigma), q_lower, q_upper, min_valid, rows_local, cols);
  ++lc.mine;
  return 0;
}
int launch_feature_extraction(const float* elev, float* const out7[7], const DeviceState* st,
                              float radius, double res, int min_valid, float p_lo, float p_hi,
                              cudaStream_t s, LaunchCounter& lc, int rows_local, int cols) {
  const int half = static_cast<int>(ceil(static_cast<double>(radius) / res));
  if (half > kMaxRegionHalf) return 1;
  const size_t n = static_cast<size_t>(rows_local) * cols;
  if (n == 0) return 0;
  FeatureLayers fl{out7[0], out7[1], out7[2], out7[3], out7[4], out7[5], out7[6]};
  // how many of the smallest / largest heights the two percentile ranks can ever reach:
  // lo = int(p_lo (count-1)) and (count-1) - int(p_hi (count-1)) are non-decreasing in count,
  // which the region's cell count bounds (+1 for float rounding)
  int region = 0;
  const float r2 = radius * radius;
  for (int dr = -half; dr <= half; ++dr)
    for (int dc = -half; dc <= half; ++dc)
      if (static_cast<float>(static_cast<double>(dr * dr + dc * dc) * res * res) <= r2) ++region;
  int keep_small = 0, keep_large = 0;
  if (region > 0) {
    const int ks = static_cast<int>(p_lo * (region - 1)) + 2;
    const int kl = (region - 1) - static_cast<int>(p_hi * (region - 1)) + 2;
    if (ks <= kSelectMax && kl <= kSelectMax) { keep_small = ks; keep_large = kl; }
  }
  feature_extraction_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, s>>>(
      elev, fl, st, half, r2, min_valid, p_lo, p_hi, rows_local, cols, keep_small, keep_large);
  ++lc.mine;
  return 0;
}
void launch_pack_pointcloud2(const float* elev, const float* const* fields, int n_fields, int rows,
                             int cols, int sub_r0, int sub_c0, int sub_rows, int sub_cols,
                             const DeviceState* st, uint32_t* col_count, uint32_t* col_offset,
                             uint32_t* total, float* out, int phase, cudaStream_t s,
                             LaunchCounter& lc) {
  PackParams p{};
  p.elev = elev;
  for (int i = 0; i < n_fields; ++i) p.field[i] = fields[i];
  p.n_fields = n_fields;
  p.rows = rows; p.cols = cols;
  p.sub_r0 = sub_r0; p.sub_c0 = sub_c0; p.sub_rows = sub_rows; p.sub_cols = sub_cols;
  const unsigned grid = static_cast<unsigned>((static_cast<size_t>(sub_cols) * 32 + kBlock - 1) / kBlock);
  if (phase == 0) {
    pack_count_kernel<<<grid, kBlock, 0, s>>>(p, col_count);
    pack_scan_kernel<<<1, 1024, 0, s>>>(col_count, col_offset, sub_cols, total);
    lc.mine += 2;
  } else {
    pack_write_kernel<<<grid, kBlock, 0, s>>>(p, st, col_offset, out);
    ++lc.mine;
  }
}
void launch_inpaint_stripe(const float* src, float* dst, int rows_local, int cols, const float* above,
                           const float* below, int min_valid, cudaStream_t s, LaunchCounter& lc) {
  inpaint_stripe_kernel<<<grid_for(static_cast<size_t>(rows_local) * cols, kBlock), kBlock, 0, s>>>(
      src, dst, rows_local, cols, above, below, min_valid);
  ++lc.mine;
}
void launch_inpaint_iter(const float* src, float* dst, const DeviceState* st, int min_valid,
                         cudaStream_t s, LaunchCounter& lc, int rows_local, int cols) {
  inpaint_iter_kernel<<<grid_for(static_cast<size_t>(rows_local) * cols, kBlock), kBlock, 0, s>>>(
      src, dst, st, min_valid, rows_local, cols);
  ++lc.mine;
}

}  // namespace fdem
