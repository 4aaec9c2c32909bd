// kernels_tile.cu — the tile path: a 2-level sort-by-cell built for this workload.
//
// A scan's points land in a few hundred "buckets" of 1024 consecutive cell keys.  Instead
// of a general radix sort of the whole scan (5 CUB launches, ~60 us at 131K points on
// B200 — launch/latency bound), the scan is sorted hierarchically:
//
//   L1  radix partition by bucket, ONE global pass:
//         K1   per-bucket point histogram (kernels.cu, fused into preprocess+bin)
//         K2   segment allocation per non-empty bucket (kernels.cu, fused into commit)
//         scatter_records_kernel: consecutive same-cell points of a warp are pre-reduced
//              with a segmented shuffle scan (LiDAR rings / image rows are coherent), and
//              one 32-byte CellRecord per run is written into its bucket's segment
//   L2  tile_estimate_kernel, one CTA per non-empty bucket (dynamic work list):
//         the bucket's records are staged into shared memory with TMA bulk copies
//         (cp.async.bulk + mbarrier), counting-sorted by cell inside shared memory, and
//         folded per cell into shared-memory accumulators — the owner thread walks a
//         cell's (few) sorted records; crowded cells get a warp and a shuffle butterfly.
//         Then the bucket's touched cells are compacted and each gets ONE Kalman / P2
//         step: every layer value is loaded once (batched) and stored once with a plain
//         store, neighbouring threads on neighbouring cells.
//
// Atomics appear only on scan-sized scratch (bucket cursors, shared-memory bins, scan
// statistics) — never on estimator state.  Results are independent of every ordering the
// atomics leave open because CellObs::combine carries explicit point-index tie-breaks.
#include <float.h>
#include <math.h>

#include "device_types.h"
#include "estimator.cuh"

namespace fdem {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 1024;  // records staged per bulk copy (32 KiB)
constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kHotCell = 48;  // records of one cell in one chunk above which a warp takes over

// ── shared-memory / TMA plumbing (inline PTX; SASS: UBLKCP + SYNCS) ──────────────
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ CellObs obs_of(const CellRecord& r) {
  CellObs o;
  o.mz = r.mz; o.mv = r.mv; o.mi = r.mi; o.xz = r.xz; o.it = r.it; o.fi = r.fi; o.li = r.li;
  return o;
}

// ───────────────────────────── L1: scatter into bucket segments ──────────────
__global__ void __launch_bounds__(kThreads)
scatter_records_kernel(const __grid_constant__ ScatterParams p) {
  pdl_launch_dependents();
  pdl_wait();  // K1 (keys, pm) and K2 (bucket segments) are complete from here on
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t INV = p.invalid_key;
  uint32_t key = INV;
  CellObs v = obs_identity();
  // all loads of this thread are independent: issue them together
  const uint32_t n_inside = p.counters[CNT_INSIDE];
  const uint32_t n_prev = p.st_cur->touched_count;
  if (i < p.n) {
    key = __ldg(&p.keys[i]);
    const float4 q = __ldg(&p.pm[i]);
    const bool has_i = p.intensity != nullptr;
    const float in = has_i ? __ldg(&p.intensity[i]) : 0.0f;
    if (key != INV) v = obs_from_point(q.z, q.w, in, has_i, i);
  }
  // updateObstacle's map_.clear(obstacle) (elevation_mapping.cpp:146) restricted to the cells
  // that can hold a value: those the last observing scan touched.  Runs only when this scan
  // has observations (update() returns early otherwise, :116-117); K3t writes this scan's
  // values afterwards.
  if (n_inside > 0 && p.obstacle) {
    for (uint32_t j = i; j < n_prev; j += gridDim.x * blockDim.x) {
      const uint32_t k = p.touched_keys[j];
      if (k != INV) p.obstacle[k] = nan_f32();
    }
  }
  // runs of consecutive lanes that hit the same cell
  uint32_t kprev = __shfl_up_sync(0xffffffffu, key, 1);
  uint32_t knext = __shfl_down_sync(0xffffffffu, key, 1);
  const bool head = lane == 0 || key != kprev;
  const bool tail = lane == 31 || key != knext;
  const uint32_t heads = __ballot_sync(0xffffffffu, head);
  const int s = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // lane 0 is always a head
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const CellObs o = obs_shfl_up(v, d);
    if (lane - d >= s) v = obs_combine(o, v);
  }
  const bool emit = tail && key != INV;
  const uint32_t emit_m = __ballot_sync(0xffffffffu, emit);
  if (emit) {
    const uint32_t bucket = key >> kBucketBits;
    // one cursor atomic per (warp, bucket); scratch, not map state
    const uint32_t peers = __match_any_sync(emit_m, bucket);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&p.tb.bucket_cursor[bucket], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t slot = p.tb.bucket_offset[bucket] + base + __popc(peers & ((1u << lane) - 1u));
    uint4* dst = reinterpret_cast<uint4*>(p.tb.records + slot);
    dst[0] = make_uint4(key & (kBucketCells - 1u), __float_as_uint(v.mz), __float_as_uint(v.mv), v.mi);
    dst[1] = make_uint4(__float_as_uint(v.xz), __float_as_uint(v.it), v.fi, v.li);
  }
}

// phase timeline of CTA 0's first bucket (SM clock ticks since kernel entry), for tuning:
// read back through fdem_mapper_debug_phase_clocks().  One thread, a dozen clock reads.
__device__ long long g_k3t_clocks[16];
// wall-clock (globaltimer, ns) entry / exit of every CTA of the last launch (first 512 CTAs)
__device__ unsigned long long g_k3t_cta_ns[2 * 512];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K3T_MARK(i) do { if (blockIdx.x == 0 && tid == 0 && first_job) g_k3t_clocks[i] = clock64() - t_entry; } while (0)

// ───────────────────────────── L2: per-bucket sort + reduce + estimate ───────
struct TileSmem {
  CellRecord stage[kChunk];          // TMA destination
  uint32_t binoff[kBucketCells];     // counting-sort bins (count -> offset -> cursor)
  float a_mz[kBucketCells];          // per-cell accumulators for the whole bucket
  float a_mv[kBucketCells];
  uint32_t a_mi[kBucketCells];
  float a_xz[kBucketCells];
  float a_it[kBucketCells];
  uint32_t a_fi[kBucketCells];
  uint32_t a_li[kBucketCells];
  uint16_t perm[kChunk];             // sorted position -> index into stage
  uint16_t tlist[kBucketCells];      // compacted list of the bucket's touched cells
  uint64_t mbar;
  uint32_t warp_sums[kWarps];
  uint32_t n_touched;
  uint32_t list_base;
  uint32_t is_last;
  uint32_t n_hot;
  uint16_t hot_cell[kChunk / kHotCell + 1];  // cells deferred to the warp-cooperative path
};

__global__ void __launch_bounds__(kThreads, 3)
tile_estimate_kernel(const __grid_constant__ EstimateParams p,
                     const __grid_constant__ TileBuffers tb, uint32_t* __restrict__ counters,
                     DeviceState* __restrict__ st_out, const __grid_constant__ PublishArgs pub) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  TileSmem& S = *reinterpret_cast<TileSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const long long t_entry = clock64();
  bool first_job = true;
  if (tid == 0 && blockIdx.x < 512) g_k3t_cta_ns[2 * blockIdx.x] = globaltimer_ns();

  // prologue that needs nothing from the scatter kernel: overlaps its tail under PDL
  if (tid == 0) mbar_init(&S.mbar, 1);
  uint32_t phase = 0;
  for (int c = tid; c < static_cast<int>(kBucketCells); c += kThreads) {
    S.a_mz[c] = FLT_MAX; S.a_mv[c] = 0.0f; S.a_mi[c] = kNone; S.a_xz[c] = -FLT_MAX;
    S.a_it[c] = -INFINITY; S.a_fi[c] = kNone; S.a_li[c] = 0u;
    S.binoff[c] = 0;
  }
  if (tid == 0) { S.n_touched = 0; S.n_hot = 0; }
  pdl_wait();  // the bucket segments are complete and visible from here on
  K3T_MARK(0);   // prologue done
  const uint32_t n_jobs = counters[CNT_BUCKETS];

  // static round-robin over the non-empty buckets K2 listed: no work-fetch atomics
  for (uint32_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    __syncthreads();  // previous bucket completely done (also publishes the mbarrier init)
    const uint4 entry = tb.bucket_list[job];  // {bucket, first record slot, points, -}
    const uint32_t b = entry.x;
    const uint32_t off = entry.y;
    // Stage the first chunk right away.  The record count is not known yet (it is being
    // loaded below); the bucket's POINT count bounds it and the segment has that many
    // slots, so copying min(points, chunk) records is in bounds and always enough.
    const uint32_t first_n = min(entry.z, static_cast<uint32_t>(kChunk));
    if (tid == 0) {
      fence_proxy_async();  // earlier generic-proxy reads of `stage` precede the async write
      mbar_expect_tx(&S.mbar, first_n * static_cast<uint32_t>(sizeof(CellRecord)));
      tma_load_1d(S.stage, tb.records + off, first_n * static_cast<uint32_t>(sizeof(CellRecord)),
                  &S.mbar);
    }
    const uint32_t nrec = tb.bucket_cursor[b];  // in flight together with the bulk copy
    K3T_MARK(1);  // job entry + record count loaded
    if (blockIdx.x == 0 && tid == 0 && first_job) g_k3t_clocks[15] = nrec;

    for (uint32_t cs = 0; cs < nrec; cs += kChunk) {
      const uint32_t cn = min(static_cast<uint32_t>(kChunk), nrec - cs);
      if (cs > 0) {
        // overflow chunks of a crowded bucket: same staging buffer, exact byte count
        if (tid == 0) {
          fence_proxy_async();
          mbar_expect_tx(&S.mbar, cn * static_cast<uint32_t>(sizeof(CellRecord)));
          tma_load_1d(S.stage, tb.records + off + cs,
                      cn * static_cast<uint32_t>(sizeof(CellRecord)), &S.mbar);
        }
        for (int c = tid; c < static_cast<int>(kBucketCells); c += kThreads) S.binoff[c] = 0;
      }
      __syncthreads();
      mbar_wait(&S.mbar, phase);
      phase ^= 1;
      if (cs == 0) K3T_MARK(2);  // first chunk staged

      // ── counting sort by cell, in shared memory ──
      for (uint32_t e = tid; e < cn; e += kThreads) atomicAdd(&S.binoff[S.stage[e].lkey], 1u);
      __syncthreads();
      {
        // exclusive scan of the 1024 bins: 4 consecutive bins per thread
        const uint32_t v0 = S.binoff[tid * 4 + 0], v1 = S.binoff[tid * 4 + 1];
        const uint32_t v2 = S.binoff[tid * 4 + 2], v3 = S.binoff[tid * 4 + 3];
        const uint32_t t = v0 + v1 + v2 + v3;
        uint32_t inc = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += o;
        }
        if (lane == 31) S.warp_sums[warp] = inc;
        __syncthreads();
        uint32_t wprefix = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
          if (w < warp) wprefix += S.warp_sums[w];
        const uint32_t base = wprefix + inc - t;
        S.binoff[tid * 4 + 0] = base;
        S.binoff[tid * 4 + 1] = base + v0;
        S.binoff[tid * 4 + 2] = base + v0 + v1;
        S.binoff[tid * 4 + 3] = base + v0 + v1 + v2;
      }
      __syncthreads();
      for (uint32_t e = tid; e < cn; e += kThreads) {
        const uint32_t pos = atomicAdd(&S.binoff[S.stage[e].lkey], 1u);
        S.perm[pos] = static_cast<uint16_t>(e);
      }
      __syncthreads();
      if (cs == 0) K3T_MARK(3);  // first chunk counting-sorted

      // ── per-cell reduction of the sorted chunk ──
      // After the counting sort the records of cell c occupy sorted positions
      // [binoff[c-1], binoff[c]).  Thread t owns cells t, t+256, t+512, t+768 of the
      // bucket: it folds each cell's few records (pre-reduced runs, typically 0-3) into
      // that cell's accumulator, which nobody else touches.  Cells with many records in
      // this chunk are deferred to the warp-cooperative path below.
      uint32_t my_hot = 0;  // bit r set: my r-th cell was deferred
#pragma unroll
      for (int r = 0; r < static_cast<int>(kBucketCells) / kThreads; ++r) {
        const int c = tid + r * kThreads;
        const uint32_t s0 = c ? S.binoff[c - 1] : 0u;
        const uint32_t e0 = S.binoff[c];
        const uint32_t cnt = e0 - s0;
        if (cnt == 0) continue;
        if (cnt > kHotCell) {
          const uint32_t h = atomicAdd(&S.n_hot, 1u);
          S.hot_cell[h] = static_cast<uint16_t>(c);
          my_hot |= 1u << r;
          continue;
        }
        CellObs a;
        a.mz = S.a_mz[c]; a.mv = S.a_mv[c]; a.mi = S.a_mi[c]; a.xz = S.a_xz[c];
        a.it = S.a_it[c]; a.fi = S.a_fi[c]; a.li = S.a_li[c];
        for (uint32_t j = s0; j < e0; ++j) a = obs_combine(a, obs_of(S.stage[S.perm[j]]));
        S.a_mz[c] = a.mz; S.a_mv[c] = a.mv; S.a_mi[c] = a.mi; S.a_xz[c] = a.xz;
        S.a_it[c] = a.it; S.a_fi[c] = a.fi; S.a_li[c] = a.li;
      }
      __syncthreads();
      const uint32_t n_hot = S.n_hot;
      if (n_hot) {
        // crowded cells: one warp per cell, lanes stride over its records, butterfly
        // reduction over warp shuffles, lane 0 merges into the accumulator
        for (uint32_t h = warp; h < n_hot; h += kWarps) {
          const int c = S.hot_cell[h];
          const uint32_t s0 = c ? S.binoff[c - 1] : 0u;
          const uint32_t e0 = S.binoff[c];
          CellObs a = obs_identity();
          for (uint32_t j = s0 + lane; j < e0; j += 32) a = obs_combine(a, obs_of(S.stage[S.perm[j]]));
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            CellObs o;
            o.mz = __shfl_xor_sync(0xffffffffu, a.mz, d);
            o.mv = __shfl_xor_sync(0xffffffffu, a.mv, d);
            o.mi = __shfl_xor_sync(0xffffffffu, a.mi, d);
            o.xz = __shfl_xor_sync(0xffffffffu, a.xz, d);
            o.it = __shfl_xor_sync(0xffffffffu, a.it, d);
            o.fi = __shfl_xor_sync(0xffffffffu, a.fi, d);
            o.li = __shfl_xor_sync(0xffffffffu, a.li, d);
            a = obs_combine(a, o);
          }
          if (lane == 0) {
            CellObs t;
            t.mz = S.a_mz[c]; t.mv = S.a_mv[c]; t.mi = S.a_mi[c]; t.xz = S.a_xz[c];
            t.it = S.a_it[c]; t.fi = S.a_fi[c]; t.li = S.a_li[c];
            t = obs_combine(t, a);
            S.a_mz[c] = t.mz; S.a_mv[c] = t.mv; S.a_mi[c] = t.mi; S.a_xz[c] = t.xz;
            S.a_it[c] = t.it; S.a_fi[c] = t.fi; S.a_li[c] = t.li;
          }
        }
        __syncthreads();
        if (tid == 0) S.n_hot = 0;
      }
      (void)my_hot;
      __syncthreads();  // accumulators + stage reads done before the next chunk / final pass
      if (cs == 0) K3T_MARK(4);  // first chunk reduced
    }
    K3T_MARK(5);  // all chunks done

    // ── compact the bucket's touched cells (ascending by construction of the ranks) ──
#pragma unroll
    for (int r = 0; r < static_cast<int>(kBucketCells) / kThreads; ++r) {
      const int c = tid + r * kThreads;
      const bool touched = S.a_fi[c] != kNone;
      const uint32_t tm = __ballot_sync(0xffffffffu, touched);
      uint32_t wbase = 0;
      if (lane == 0 && tm) wbase = atomicAdd(&S.n_touched, __popc(tm));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (touched) S.tlist[wbase + __popc(tm & ((1u << lane) - 1u))] = static_cast<uint16_t>(c);
    }
    __syncthreads();
    const uint32_t nt = S.n_touched;
    K3T_MARK(6);  // touched cells compacted
    if (tid == 0) {
      // reserve this bucket's slice of the touched-cell list (next scan's obstacle reset)
      // and count the cells; the round trip overlaps the estimator work below
      S.list_base = nt ? atomicAdd(&st_out->touched_count, nt) : 0u;
      if (nt) atomicAdd(&counters[CNT_CELLS], nt);
      tb.bucket_count[b] = 0;   // re-arm the L1 scratch for the next scan
      tb.bucket_cursor[b] = 0;
    }

    // ── estimator: one Kalman / P2 step per touched cell; each layer value is loaded once
    //    (batched) and stored once; neighbouring threads handle neighbouring cells ──
    const uint32_t key_base = b << kBucketBits;
    for (uint32_t t = tid; t < nt; t += kThreads) {
      const uint32_t c = S.tlist[t];
      CellObs v;
      v.mz = S.a_mz[c]; v.mv = S.a_mv[c]; v.mi = S.a_mi[c]; v.xz = S.a_xz[c];
      v.it = S.a_it[c]; v.fi = S.a_fi[c]; v.li = S.a_li[c];
      apply_observation(p, key_base + c, v);
    }
    __syncthreads();
    K3T_MARK(7);  // estimator done
    const uint32_t lb = S.list_base;
    for (uint32_t t = tid; t < nt; t += kThreads) {
      const uint32_t c = S.tlist[t];
      p.touched_keys[lb + t] = key_base + c;
      if (p.touched_minz) p.touched_minz[lb + t] = S.a_mz[c];
    }
    K3T_MARK(8);  // touched list written
    first_job = false;
    if (job + gridDim.x < n_jobs) {
      // another bucket follows on this CTA: re-arm the accumulators and bins
      __syncthreads();
      for (int c = tid; c < static_cast<int>(kBucketCells); c += kThreads) {
        S.a_mz[c] = FLT_MAX; S.a_mv[c] = 0.0f; S.a_mi[c] = kNone; S.a_xz[c] = -FLT_MAX;
        S.a_it[c] = -INFINITY; S.a_fi[c] = kNone; S.a_li[c] = 0u;
        S.binoff[c] = 0;
      }
      if (tid == 0) S.n_touched = 0;
    }
  }

  // ── end of scan: the LAST CTA to finish publishes (no extra kernel / memset / memcpy):
  // scan statistics + committed state go to the host through mapped pinned memory, the
  // committed state becomes current, and the counters are re-armed for the next scan ──
  first_job = true;
  K3T_MARK(9);  // all jobs of CTA 0 done
  if (pub.enabled) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      S.is_last = (atomicAdd(&counters[CNT_WORK], 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (S.is_last && warp == 0) {
      __threadfence();
      constexpr int kStateWords = sizeof(DeviceState) / 4;
      static_assert(kStateWords <= 32 && CNT_COUNT <= 32, "one warp publishes");
      volatile uint32_t* vc = counters;
      volatile const uint32_t* vs = reinterpret_cast<const uint32_t*>(st_out);
      uint32_t c = 0, w = 0;
      if (lane < CNT_COUNT) c = vc[lane];
      if (lane < kStateWords) w = vs[lane];
      if (lane < CNT_COUNT) {
        pub.host_out[lane] = c;
        counters[lane] = 0;
      }
      if (lane < kStateWords) {
        pub.host_out[CNT_COUNT + lane] = w;
        reinterpret_cast<uint32_t*>(pub.st_cur)[lane] = w;
      }
    }
  }
  K3T_MARK(10);  // kernel exit of CTA 0
  if (tid == 0 && blockIdx.x < 512) g_k3t_cta_ns[2 * blockIdx.x + 1] = globaltimer_ns();
}

}  // namespace

int tile_estimate_debug_cta_ns(unsigned long long* out1024) {
  return static_cast<int>(cudaMemcpyFromSymbol(out1024, g_k3t_cta_ns, sizeof(unsigned long long) * 1024));
}

int tile_estimate_debug_clocks(long long* out16) {
  return static_cast<int>(cudaMemcpyFromSymbol(out16, g_k3t_clocks, sizeof(long long) * 16));
}

int tile_estimate_configure() {
  return static_cast<int>(cudaFuncSetAttribute(tile_estimate_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(sizeof(TileSmem))));
}

void launch_scatter_records(const ScatterParams& p, cudaStream_t s, LaunchCounter& lc) {
  if (p.n == 0) return;
  scatter_records_kernel<<<(p.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(p);
  ++lc.mine;
}

void launch_tile_estimate(const EstimateParams& p, const TileBuffers& tb, uint32_t* counters,
                          DeviceState* st_out, const PublishArgs& pub, cudaStream_t s,
                          LaunchCounter& lc) {
  // CTAs stride over the non-empty-bucket list; 3 CTAs/SM fit in shared memory
  const uint32_t grid = min(tb.n_buckets, 148u * 3u);
  tile_estimate_kernel<<<grid, kThreads, sizeof(TileSmem), s>>>(p, tb, counters, st_out, pub);
  ++lc.mine;
}

}  // namespace fdem

namespace fdem {
KernelDesc desc_scatter_records(uint32_t n) {
  return KernelDesc{reinterpret_cast<const void*>(&scatter_records_kernel),
                    dim3((n + kThreads - 1) / kThreads), dim3(kThreads), 0};
}
KernelDesc desc_tile_estimate(uint32_t n_buckets) {
  return KernelDesc{reinterpret_cast<const void*>(&tile_estimate_kernel),
                    dim3(n_buckets < 148u * 3u ? n_buckets : 148u * 3u), dim3(kThreads),
                    sizeof(TileSmem)};
}
}  // namespace fdem
