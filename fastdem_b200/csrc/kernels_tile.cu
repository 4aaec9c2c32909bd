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
//   L2  tile_estimate_kernel, one CTA per non-empty bucket (static round-robin work list):
//         the bucket's records are staged into shared memory with TMA bulk copies
//         (cp.async.bulk + mbarrier), counting-sorted by cell inside shared memory, and
//         reduced per cell — the owner thread walks a cell's (few) sorted records; crowded
//         cells get a warp and a shuffle butterfly.  A bucket whose records fit one chunk
//         (the common case) is reduced straight into a compact list of touched cells; a
//         bucket that needs several chunks goes through per-cell accumulators.  Each
//         touched cell then gets ONE Kalman / P2 step: every layer value is loaded once
//         (batched) and stored once with a plain store, neighbouring threads on
//         neighbouring cells.
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

constexpr int kThreads = 256;  // scatter kernel
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

// system-scope flag traffic of the multi-GPU handshake (flags live in peer-mapped memory)
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ───────────────────────────── L1: scatter into bucket segments ──────────────
__device__ __forceinline__ void scatter_records_body(const ScatterParams& p, uint32_t n_pts, uint32_t index_base) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t INV = p.invalid_key;
  uint32_t key = INV;
  CellObs v = obs_identity();
  // all loads of this thread are independent: issue them together
  const uint32_t n_inside = p.counters[CNT_INSIDE];
  const uint32_t n_prev = p.st_cur->touched_count;
  if (i < n_pts) {
    key = __ldg(&p.keys[i]);
    const float4 q = __ldg(&p.pm[i]);
    const bool has_i = p.intensity != nullptr;
    const float in = has_i ? __ldg(&p.intensity[p.slice ? index_base + i : i]) : 0.0f;
    if (key != INV) v = obs_from_point(q.z, q.w, in, has_i, i + index_base);
  }
  // updateObstacle's map_.clear(obstacle) (elevation_mapping.cpp:146) restricted to the cells
  // that can hold a value: those the last observing scan touched (the whole layer when the
  // caller edited it, SF_OBSTACLE_DIRTY).  Runs only when this scan has observations (update()
  // returns early otherwise, :116-117); K3t writes this scan's values afterwards.
  if (n_inside > 0 && p.obstacle) {
    const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
    if (p.st_cur->flags & SF_OBSTACLE_DIRTY) {
      for (size_t c = i; c < p.obstacle_cells; c += nthreads) p.obstacle[c] = nan_f32();
    } else {
      for (size_t j = i; j < n_prev; j += nthreads) {
        const uint32_t k = p.touched_keys[j];
        if (k != INV) p.obstacle[k] = nan_f32();
      }
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
    const uint32_t bucket = key >> p.tb.bucket_bits;
    // one cursor atomic per (warp, bucket); scratch, not map state
    const uint32_t peers = __match_any_sync(emit_m, bucket);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&p.tb.bucket_cursor[bucket], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t slot = p.tb.bucket_offset[bucket] + base + __popc(peers & ((1u << lane) - 1u));
    CellRecord* area = p.owner_bps ? p.owner_records[bucket / p.owner_bps] : p.tb.records;
    uint4* dst = reinterpret_cast<uint4*>(area + slot);
    dst[0] = make_uint4(key & ((1u << p.tb.bucket_bits) - 1u), __float_as_uint(v.mz),
                        __float_as_uint(v.mv), v.mi);
    dst[1] = make_uint4(__float_as_uint(v.xz), __float_as_uint(v.it), v.fi, v.li);
  }
}

__global__ void __launch_bounds__(kThreads)
scatter_records_kernel(const __grid_constant__ ScatterParams p) {
  pdl_launch_dependents();
  pdl_wait();  // K1 (keys, pm) and K2 (bucket segments) are complete from here on
  if (!p.slice) {
    scatter_records_body(p, p.n, p.index_base);
    return;
  }
  // ── multi-GPU front half: the slice K1 binned (decided on the device); the grid covers the
  //    whole scan and the CTAs beyond the slice have nothing to scatter ──
  const uint32_t n_pts = p.slice->count;
  if (blockIdx.x * blockDim.x < n_pts) scatter_records_body(p, n_pts, p.slice->begin);
}

// Tuning probes, compiled only with -DFDEM_PROBES (tools/phase_probe.py): phase timeline of
// CTA 0's first bucket (SM clock ticks since kernel entry) and the wall-clock entry / exit of
// every CTA of the last launch.  The production build carries none of this.
#ifdef FDEM_PROBES
__device__ long long g_k3t_clocks[16];
__device__ unsigned long long g_k3t_cta_ns[2 * 512];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K3T_MARK(i) do { if (blockIdx.x == 0 && tid == 0 && first_job) g_k3t_clocks[i] = clock64() - t_entry; } while (0)
#else
#define K3T_MARK(i) do { } while (0)
#endif

// ───────────────────────────── L2: per-bucket sort + reduce + estimate ───────
// Compiled for three bucket shapes; the mapper picks one per scan (capi.cu):
//   BITS 10: 1024 cells, 256 threads (4 cells/thread), 1024-record chunks, ~68 KiB smem, 3 CTAs/SM
//            — the default: a sparse scan (a LiDAR sweep touches ~15 % of the cells) fills
//            few hundred buckets, one wave of CTAs
//   BITS  9:  512 cells, 128 threads (4 cells/thread),  512-record chunks, ~34 KiB smem, 6 CTAs/SM
//   BITS  8:  256 cells, 256 threads (1 cell/thread),  1024-record chunks, ~43 KiB smem, 3 CTAs/SM
//            — dense scans (an RGB-D frame puts ~15 points on every touched cell of a small
//            patch): parallelism has to come from the records, not from the map area
template <int BITS>
struct TileCfg {
  static constexpr int kCells = 1 << BITS;
  static constexpr int kThr = BITS == 9 ? 128 : 256;
  static constexpr int kCellsPerThread = kCells / kThr;
  static constexpr int kWarps = kThr / 32;
  static constexpr int kChunk = BITS == 9 ? 512 : 1024;  // records staged per bulk copy (32 B each)
  static constexpr int kMinBlocks = BITS == 9 ? 6 : 3;
};

template <int BITS>
struct TileSmem {
  using C = TileCfg<BITS>;
  CellRecord stage[C::kChunk];       // TMA destination
  uint32_t binoff[C::kCells];        // counting-sort bins (count -> offset -> cursor)
  // reduced observations: slot = position in the touched list (single-chunk bucket) or
  // the cell itself (multi-chunk bucket, used as accumulators)
  float a_mz[C::kCells];
  float a_mv[C::kCells];
  uint32_t a_mi[C::kCells];
  float a_xz[C::kCells];
  float a_it[C::kCells];
  uint32_t a_fi[C::kCells];
  uint32_t a_li[C::kCells];
  uint16_t perm[C::kChunk];          // sorted position -> index into stage
  uint16_t tlist[C::kCells];         // compacted list of the bucket's touched cells
  uint64_t mbar;
  uint32_t warp_sums[C::kWarps];
  uint32_t n_touched;
  uint32_t list_base;
  uint32_t is_last;
  uint32_t n_hot;
  uint16_t hot_cell[C::kChunk / kHotCell + 1];  // cells deferred to the warp-cooperative path
  uint16_t hot_slot[C::kChunk / kHotCell + 1];  // where their result goes
};

template <typename S_>
__device__ __forceinline__ CellObs acc_load(const S_& S, uint32_t a) {
  CellObs t;
  t.mz = S.a_mz[a]; t.mv = S.a_mv[a]; t.mi = S.a_mi[a]; t.xz = S.a_xz[a];
  t.it = S.a_it[a]; t.fi = S.a_fi[a]; t.li = S.a_li[a];
  return t;
}
template <typename S_>
__device__ __forceinline__ void acc_store(S_& S, uint32_t a, const CellObs& t) {
  S.a_mz[a] = t.mz; S.a_mv[a] = t.mv; S.a_mi[a] = t.mi; S.a_xz[a] = t.xz;
  S.a_it[a] = t.it; S.a_fi[a] = t.fi; S.a_li[a] = t.li;
}

// stage records [cs, cs + cn) of a sharded bucket — the concatenation of its pieces in the
// sources' record buffers (peer memory, read over NVLink by the TMA engine) — one bulk copy per
// piece that overlaps the chunk
__device__ __forceinline__ void tma_load_pieces(CellRecord* stage, const ShardJob& J, const ShardBackArgs& sh,
                                                uint32_t cs, uint32_t cn, uint64_t* bar) {
  uint32_t pos = 0;
  for (int s = 0; s < sh.world; ++s) {
    const uint32_t lo = max(cs, pos), hi = min(cs + cn, pos + J.cnt[s]);
    if (lo < hi)
      tma_load_1d(stage + (lo - cs), sh.peer_records[s] + J.off[s] + (lo - pos),
                  (hi - lo) * static_cast<uint32_t>(sizeof(CellRecord)), bar);
    pos += J.cnt[s];
  }
}

// SHARD = false: one GPU, jobs = the bucket list K2 wrote, records in this GPU's scratch.
// SHARD = true: multi-GPU GLOBAL map, back half: jobs = the ShardJob list shard_gather_kernel
// built for this rank's stripe, records pulled from every source rank's arena.
template <int BITS, bool SHARD>
__device__ __forceinline__ void
tile_estimate_body(const EstimateParams& p, const TileBuffers& tb, uint32_t* __restrict__ counters,
                   DeviceState* __restrict__ st_out, const PublishArgs& pub, const ShardBackArgs* shp) {
  using C = TileCfg<BITS>;
  constexpr int kThr = C::kThr;
  constexpr int kCells = C::kCells;
  constexpr int kChunk = C::kChunk;
  constexpr int kWarps = C::kWarps;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  TileSmem<BITS>& S = *reinterpret_cast<TileSmem<BITS>*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
#ifdef FDEM_PROBES
  const long long t_entry = clock64();
  bool first_job = true;
  if (tid == 0 && blockIdx.x < 512) g_k3t_cta_ns[2 * blockIdx.x] = globaltimer_ns();
#endif

  // batched graphs: the next scan's back prologue may launch now and wait for this grid to finish
  pdl_launch_dependents();
  // prologue that needs nothing from the scatter kernel: overlaps its tail under PDL
  if (tid == 0) mbar_init(&S.mbar, 1);
  uint32_t phase = 0;
  for (int c = tid; c < kCells; c += kThr) S.binoff[c] = 0;
  if (tid == 0) { S.n_touched = 0; S.n_hot = 0; }
  pdl_wait();  // the bucket segments are complete and visible from here on
  K3T_MARK(0);   // prologue done
  // the job count and this CTA's first list entry are independent loads: one round trip
  // (the list has one slot per bucket and the grid never exceeds that, so the speculative
  // read is in bounds; it is used only when job < n_jobs, i.e. when this scan wrote it)
  // the work list: K2's bucket list, or what the light pass left over
  const bool leftovers = (SHARD ? shp->job_counter : tb.job_counter) == CNT_HEAVY;
  const uint4* __restrict__ job_list = leftovers ? tb.heavy_list : tb.bucket_list;
  const ShardJob* __restrict__ shard_jobs = SHARD ? (leftovers ? shp->heavy_jobs : shp->jobs) : nullptr;
  const uint32_t n_jobs = counters[leftovers ? CNT_HEAVY : CNT_BUCKETS];
  uint4 entry_next = make_uint4(0u, 0u, 0u, 0u);
  if (!SHARD) entry_next = job_list[blockIdx.x];

  // static round-robin over the non-empty buckets K2 listed: no work-fetch atomics
  for (uint32_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    __syncthreads();  // previous bucket completely done (also publishes the mbarrier init)
    uint32_t b, off = 0, nrec;
    ShardJob J;
    if (SHARD) {
      J = shard_jobs[job];  // {bucket, records, per-source (first slot, count)}
      b = J.bucket;
      nrec = J.total;
      if (tid == 0) {
        const uint32_t first_n = min(nrec, static_cast<uint32_t>(kChunk));
        fence_proxy_async();
        mbar_expect_tx(&S.mbar, first_n * static_cast<uint32_t>(sizeof(CellRecord)));
        tma_load_pieces(S.stage, J, *shp, 0u, first_n, &S.mbar);
      }
    } else {
      const uint4 entry = entry_next;  // {bucket, first record slot, points, -}
      if (job + gridDim.x < n_jobs) entry_next = job_list[job + gridDim.x];  // prefetch
      b = entry.x;
      off = entry.y;
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
      nrec = tb.bucket_cursor[b];  // in flight together with the bulk copy
    }
    K3T_MARK(1);  // job entry + record count loaded
#ifdef FDEM_PROBES
    if (blockIdx.x == 0 && tid == 0 && first_job) g_k3t_clocks[15] = nrec;
#endif
    // single-chunk bucket: cells are reduced from scratch straight into the touched list;
    // multi-chunk bucket: per-cell accumulators carry a cell across chunks
    const bool single = nrec <= static_cast<uint32_t>(kChunk);
    if (!single) {
      for (int c = tid; c < kCells; c += kThr) acc_store(S, c, obs_identity());
    }

    for (uint32_t cs = 0; cs < nrec; cs += kChunk) {
      const uint32_t cn = min(static_cast<uint32_t>(kChunk), nrec - cs);
      if (cs > 0) {
        // overflow chunks of a crowded bucket: same staging buffer, exact byte count
        if (tid == 0) {
          fence_proxy_async();
          mbar_expect_tx(&S.mbar, cn * static_cast<uint32_t>(sizeof(CellRecord)));
          if (SHARD) tma_load_pieces(S.stage, J, *shp, cs, cn, &S.mbar);
          else tma_load_1d(S.stage, tb.records + off + cs,
                           cn * static_cast<uint32_t>(sizeof(CellRecord)), &S.mbar);
        }
        for (int c = tid; c < kCells; c += kThr) S.binoff[c] = 0;
      }
      __syncthreads();
      mbar_wait(&S.mbar, phase);
      phase ^= 1;
      if (cs == 0) K3T_MARK(2);  // first chunk staged

      // ── counting sort by cell, in shared memory ──
      for (uint32_t e = tid; e < cn; e += kThr) atomicAdd(&S.binoff[S.stage[e].lkey], 1u);
      __syncthreads();
      {
        // exclusive scan of the bins: kCellsPerThread consecutive bins per thread
        constexpr int CPT = C::kCellsPerThread;
        uint32_t v[CPT];
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          v[k] = S.binoff[tid * CPT + k];
          t += v[k];
        }
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
        uint32_t base = wprefix + inc - t;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          S.binoff[tid * CPT + k] = base;
          base += v[k];
        }
      }
      __syncthreads();
      for (uint32_t e = tid; e < cn; e += kThr) {
        const uint32_t pos = atomicAdd(&S.binoff[S.stage[e].lkey], 1u);
        S.perm[pos] = static_cast<uint16_t>(e);
      }
      __syncthreads();
      if (cs == 0) K3T_MARK(3);  // first chunk counting-sorted

      // ── per-cell reduction of the sorted chunk ──
      // After the counting sort the records of cell c occupy sorted positions
      // [binoff[c-1], binoff[c]).  Thread t owns cells t, t+kThr, ... of the bucket and folds
      // each cell's few records (pre-reduced runs, typically 0-3).  Cells
      // with many records in this chunk are deferred to the warp-cooperative path below.
#pragma unroll
      for (int r = 0; r < kCells / kThr; ++r) {
        const int c = tid + r * kThr;
        const uint32_t s0 = c ? S.binoff[c - 1] : 0u;
        const uint32_t e0 = S.binoff[c];
        const uint32_t cnt = e0 - s0;
        uint32_t slot = static_cast<uint32_t>(c);
        if (single) {
          // the cell's slot in the touched list (the loop is warp-uniform up to here)
          const uint32_t tm = __ballot_sync(0xffffffffu, cnt != 0);
          uint32_t wbase = 0;
          if (lane == 0 && tm) wbase = atomicAdd(&S.n_touched, __popc(tm));
          wbase = __shfl_sync(0xffffffffu, wbase, 0);
          slot = wbase + __popc(tm & ((1u << lane) - 1u));
          if (cnt) S.tlist[slot] = static_cast<uint16_t>(c);
        }
        if (cnt == 0) continue;
        if (cnt > kHotCell) {
          const uint32_t h = atomicAdd(&S.n_hot, 1u);
          S.hot_cell[h] = static_cast<uint16_t>(c);
          S.hot_slot[h] = static_cast<uint16_t>(slot);
          continue;
        }
        CellObs a = single ? obs_identity() : acc_load(S, slot);
        for (uint32_t j = s0; j < e0; ++j) a = obs_combine(a, obs_of(S.stage[S.perm[j]]));
        acc_store(S, slot, a);
      }
      __syncthreads();
      const uint32_t n_hot = S.n_hot;
      if (n_hot) {
        // crowded cells: one warp per cell, lanes stride over its records, butterfly
        // reduction over warp shuffles, lane 0 merges into the cell's slot
        for (uint32_t h = warp; h < n_hot; h += kWarps) {
          const int c = S.hot_cell[h];
          const uint32_t slot = S.hot_slot[h];
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
            if (!single) a = obs_combine(acc_load(S, slot), a);
            acc_store(S, slot, a);
          }
        }
        __syncthreads();
        if (tid == 0) S.n_hot = 0;
      }
      __syncthreads();  // results + stage reads done before the next chunk / final pass
      if (cs == 0) K3T_MARK(4);  // first chunk reduced
    }
    K3T_MARK(5);  // all chunks done

    if (!single) {
      // ── compact the bucket's touched cells out of the accumulators ──
#pragma unroll
      for (int r = 0; r < kCells / kThr; ++r) {
        const int c = tid + r * kThr;
        const bool touched = S.a_fi[c] != kNone;
        const uint32_t tm = __ballot_sync(0xffffffffu, touched);
        uint32_t wbase = 0;
        if (lane == 0 && tm) wbase = atomicAdd(&S.n_touched, __popc(tm));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (touched) S.tlist[wbase + __popc(tm & ((1u << lane) - 1u))] = static_cast<uint16_t>(c);
      }
      __syncthreads();
    }
    const uint32_t nt = S.n_touched;
    K3T_MARK(6);  // touched cells compacted
    if (tid == kThr - 1) {
      // reserve this bucket's slice of the touched-cell list (next scan's obstacle reset)
      // and count the cells.  Done by the thread least likely to own an estimator step, so
      // the atomic's round trip does not sit in front of a cell's loads.
      S.list_base = nt ? atomicAdd(&st_out->touched_count, nt) : 0u;
      if (nt) atomicAdd(&counters[CNT_CELLS], nt);
      if (!SHARD) {             // (sharded: the source re-arms its own tables, shard_begin_kernel)
        tb.bucket_count[b] = 0;   // re-arm the L1 scratch for the next scan
        tb.bucket_cursor[b] = 0;
      }
    }

    // ── estimator: one Kalman / P2 step per touched cell; each layer value is loaded once
    //    (batched) and stored once; neighbouring threads handle neighbouring cells ──
    const uint32_t key_base = b << BITS;
    for (uint32_t t = tid; t < nt; t += kThr) {
      const uint32_t c = S.tlist[t];
      const CellObs v = acc_load(S, single ? t : c);
      apply_observation(p, key_base + c, v);
    }
    __syncthreads();
    K3T_MARK(7);  // estimator done
    const uint32_t lb = S.list_base;
    for (uint32_t t = tid; t < nt; t += kThr) {
      const uint32_t c = S.tlist[t];
      p.touched_keys[lb + t] = key_base + c;
      if (p.touched_minz) p.touched_minz[lb + t] = S.a_mz[single ? t : c];
    }
    K3T_MARK(8);  // touched list written
#ifdef FDEM_PROBES
    first_job = false;
#endif
    if (job + gridDim.x < n_jobs) {
      // another bucket follows on this CTA: re-arm the bins
      __syncthreads();
      for (int c = tid; c < kCells; c += kThr) S.binoff[c] = 0;
      if (tid == 0) S.n_touched = 0;
    }
  }

  // ── end of scan: the LAST CTA to finish publishes (no extra kernel / memset / memcpy):
  // scan statistics + committed state go to the host through mapped pinned memory, the
  // committed state becomes current, and the counters are re-armed for the next scan ──
#ifdef FDEM_PROBES
  first_job = true;
#endif
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
      const uint32_t cells = __shfl_sync(0xffffffffu, c, CNT_CELLS);
      if (SHARD && lane < shp->world) {
        // every record of this scan has been read: the sources may reuse the buffers; with the
        // flag goes this stripe's load, which the slice split of scan seq + 2 is derived from
        shp->peer_hdr[lane]->load[shp->seq % kShardDepth][shp->rank] = cells;
        __threadfence_system();
        st_release_sys(&shp->peer_hdr[lane]->consumed[shp->rank], shp->seq);
      }
    }
  }
  K3T_MARK(10);  // kernel exit of CTA 0
#ifdef FDEM_PROBES
  if (tid == 0 && blockIdx.x < 512) g_k3t_cta_ns[2 * blockIdx.x + 1] = globaltimer_ns();
#endif
}

// ───────────────────────────── K3t, light pass: one WARP per bucket ──────────────────────
// A scan spread over a large map (C4, C5, every stripe of the multi-GPU map) fills thousands of
// buckets with a few dozen cells each.  A CTA per such bucket pays its ~7 us chain of dependent
// phases — each fenced by a block barrier — for a warp's worth of work, three buckets per SM at a
// time.  Here a bucket with <= kLightRecs records is one WARP's job: its records go to the warp's
// private shared-memory stage, a counting sort by cell and the per-cell fold run warp-
// synchronously (no block barrier), and each lane that heads a cell's run does that cell's Kalman
// / P2 step.  Eight buckets per CTA, ~24 per SM in flight.  Buckets with more records are handed
// to the CTA-per-bucket kernel through the heavy list.
constexpr int kLightRecs = 128;
constexpr int kLightWarps = 8;
struct LightWarpSmem {
  CellRecord stage[kLightRecs];
  uint32_t bins[1024];          // count -> exclusive offset -> cursor, per cell of the bucket
  uint8_t perm[kLightRecs];     // sorted position -> index into stage
  uint8_t hpos[kLightRecs];     // touched cell t -> sorted position of its first record
};

constexpr size_t kLightSmem = sizeof(LightWarpSmem) * kLightWarps;
constexpr unsigned kLightGrid = 148u * 3u;

template <bool SHARD>
__device__ __forceinline__ void
tile_estimate_light_body(const EstimateParams& p, const TileBuffers& tb, uint32_t* __restrict__ counters,
                         DeviceState* __restrict__ st_out, const ShardBackArgs* shp) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  LightWarpSmem& W = reinterpret_cast<LightWarpSmem*>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  constexpr int kRounds = kLightRecs / 32;
  pdl_launch_dependents();
  for (int i = lane; i < 1024; i += 32) W.bins[i] = 0u;
  __syncwarp();
  pdl_wait();  // the bucket segments (or the shard job list) are complete from here on
  const uint32_t n_jobs = counters[CNT_BUCKETS];
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  uint32_t cells_done = 0;
  for (uint32_t job = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; job < n_jobs; job += n_warps) {
    uint32_t b, nrec;
    uint4 entry = make_uint4(0u, 0u, 0u, 0u);
    ShardJob J;
    // ── stage the records: coalesced 32-byte reads.  One GPU: the record count is still being
    //    loaded, the bucket's POINT count bounds it and the segment has that many slots, so
    //    min(points, 128) records are read speculatively beside it (one round trip less) ──
    uint4 lo[kRounds], hi[kRounds];
    if (SHARD) {
      J = shp->jobs[job];
      b = J.bucket;
      nrec = J.total;
      if (nrec <= static_cast<uint32_t>(kLightRecs)) {
#pragma unroll
        for (int k = 0; k < kRounds; ++k) {
          const uint32_t r = lane + 32u * k;
          if (r < nrec) {
            // a sharded bucket's records lie in up to `world` pieces, local or in a peer's arena
            const CellRecord* src = nullptr;
            uint32_t pos = 0;
#pragma unroll
            for (int sidx = 0; sidx < kMaxShards; ++sidx) {
              const uint32_t c = J.cnt[sidx];
              if (src == nullptr && r < pos + c) src = shp->peer_records[sidx] + J.off[sidx] + (r - pos);
              pos += c;
            }
            lo[k] = __ldcg(reinterpret_cast<const uint4*>(src));
            hi[k] = __ldcg(reinterpret_cast<const uint4*>(src) + 1);
          }
        }
      }
    } else {
      entry = tb.bucket_list[job];   // {bucket, first record slot, points, -}
      b = entry.x;
      const uint32_t spec = min(entry.z, static_cast<uint32_t>(kLightRecs));
#pragma unroll
      for (int k = 0; k < kRounds; ++k) {
        const uint32_t r = lane + 32u * k;
        if (r < spec) {
          const uint4* src = reinterpret_cast<const uint4*>(tb.records + entry.y + r);
          lo[k] = __ldcg(src);
          hi[k] = __ldcg(src + 1);
        }
      }
      nrec = tb.bucket_cursor[b];
    }
    if (nrec > static_cast<uint32_t>(kLightRecs)) {   // warp-uniform: the CTA-per-bucket kernel takes it
      if (lane == 0) {
        const uint32_t h = atomicAdd(&counters[CNT_HEAVY], 1u);
        if (SHARD) shp->heavy_jobs[h] = J;
        else tb.heavy_list[h] = entry;
      }
      continue;
    }
    // ── counting sort by cell, warp-synchronous ──
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
      const uint32_t r = lane + 32u * k;
      if (r < nrec) {
        reinterpret_cast<uint4*>(&W.stage[r])[0] = lo[k];
        reinterpret_cast<uint4*>(&W.stage[r])[1] = hi[k];
        atomicAdd(&W.bins[lo[k].x], 1u);   // .x is the record's cell within the bucket
      }
    }
    __syncwarp();
    {
      // exclusive scan over the 1024 bins: 32 consecutive bins per lane
      uint4* my = reinterpret_cast<uint4*>(&W.bins[lane * 32]);
      uint4 v[8];
      uint32_t sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        v[q] = my[q];
        sum += v[q].x + v[q].y + v[q].z + v[q].w;
      }
      uint32_t inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
      }
      if (sum) {
        // only the touched cells' bins get their offset: the bins of empty cells stay zero for
        // good, so re-arming after the bucket costs one store per touched cell
        uint32_t base = inc - sum;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 e;
          e.x = v[q].x ? base : 0u; base += v[q].x;
          e.y = v[q].y ? base : 0u; base += v[q].y;
          e.z = v[q].z ? base : 0u; base += v[q].z;
          e.w = v[q].w ? base : 0u; base += v[q].w;
          my[q] = e;
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
      const uint32_t r = lane + 32u * k;
      if (r < nrec) W.perm[atomicAdd(&W.bins[lo[k].x], 1u)] = static_cast<uint8_t>(r);
    }
    __syncwarp();
    // ── the bucket's touched cells = the heads of the runs of equal cells in sorted order ──
    uint32_t nt = 0;
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
      const uint32_t pp = lane + 32u * k;
      bool head = false;
      if (pp < nrec) {
        const uint32_t key = W.stage[W.perm[pp]].lkey;
        head = pp == 0 || W.stage[W.perm[pp - 1]].lkey != key;
      }
      const uint32_t hm = __ballot_sync(0xffffffffu, head);
      if (head) W.hpos[nt + __popc(hm & ((1u << lane) - 1u))] = static_cast<uint8_t>(pp);
      nt += __popc(hm);
    }
    // this bucket's slice of the touched-cell list: ONE atomic per bucket, in flight while the
    // cells are folded
    uint32_t lb = 0;
    if (lane == 0) lb = atomicAdd(&st_out->touched_count, nt);
    __syncwarp();
    // re-arm the bins for this warp's next bucket (only the touched cells' bins are non-zero)
    for (uint32_t t = lane; t < nt; t += 32) W.bins[W.stage[W.perm[W.hpos[t]]].lkey] = 0u;
    if (!SHARD && lane == 0) {   // re-arm the L1 scratch for the next scan
      tb.bucket_count[b] = 0;
      tb.bucket_cursor[b] = 0;
    }
    // ── one lane per touched cell: fold its run, then that cell's estimator step ──
    lb = __shfl_sync(0xffffffffu, lb, 0);
    const uint32_t key_base = b << 10;
    for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
      const uint32_t t = t0 + lane;
      if (t < nt) {
        const uint32_t s0 = W.hpos[t];
        const uint32_t e0 = t + 1 < nt ? static_cast<uint32_t>(W.hpos[t + 1]) : nrec;
        const uint32_t key = W.stage[W.perm[s0]].lkey;
        CellObs a = obs_of(W.stage[W.perm[s0]]);
        for (uint32_t q = s0 + 1; q < e0; ++q) a = obs_combine(a, obs_of(W.stage[W.perm[q]]));
        apply_observation(p, key_base + key, a);
        p.touched_keys[lb + t] = key_base + key;
        if (p.touched_minz) p.touched_minz[lb + t] = a.mz;
      }
    }
    cells_done += nt;
    __syncwarp();   // every read of this bucket's stage / perm / hpos precedes the next bucket's writes
  }
  if (lane == 0 && cells_done) atomicAdd(&counters[CNT_CELLS], cells_done);
}

__global__ void __launch_bounds__(kLightWarps * 32, 3)
tile_estimate_light_kernel(const __grid_constant__ EstimateParams p, const __grid_constant__ TileBuffers tb,
                           uint32_t* __restrict__ counters, DeviceState* __restrict__ st_out) {
  tile_estimate_light_body<false>(p, tb, counters, st_out, nullptr);
}
__global__ void __launch_bounds__(kLightWarps * 32, 3)
tile_estimate_light_shard_kernel(const __grid_constant__ EstimateParams p, const __grid_constant__ ShardBackArgs sh,
                                 uint32_t* __restrict__ counters, DeviceState* __restrict__ st_out) {
  const TileBuffers none{};
  tile_estimate_light_body<true>(p, none, counters, st_out, &sh);
}

template <int BITS>
__global__ void __launch_bounds__(TileCfg<BITS>::kThr, TileCfg<BITS>::kMinBlocks)
tile_estimate_kernel(const __grid_constant__ EstimateParams p,
                     const __grid_constant__ TileBuffers tb, uint32_t* __restrict__ counters,
                     DeviceState* __restrict__ st_out, const __grid_constant__ PublishArgs pub) {
  tile_estimate_body<BITS, false>(p, tb, counters, st_out, pub, nullptr);
}

__global__ void __launch_bounds__(TileCfg<10>::kThr, TileCfg<10>::kMinBlocks)
tile_estimate_shard_kernel(const __grid_constant__ EstimateParams p,
                           const __grid_constant__ ShardBackArgs sh, uint32_t* __restrict__ counters,
                           DeviceState* __restrict__ st_out, const __grid_constant__ PublishArgs pub) {
  const TileBuffers none{};
  tile_estimate_body<10, true>(p, none, counters, st_out, pub, &sh);
}

// ───────────────────────────── multi-GPU GLOBAL map: the handshake kernels ─────────────
// FRONT, first kernel of a scan on every rank: wait until every owner has finished reading the
// buffers this scan is about to overwrite (scan seq - 2 used the same parity), then re-arm the
// bucket tables and the scan counters.
__global__ void __launch_bounds__(256)
shard_begin_kernel(ShardHeader* __restrict__ hdr, uint32_t seq, int world, int rank, uint32_t n_scan,
                   uint32_t back_weight, ShardSlice* __restrict__ slice_out, uint32_t* __restrict__ counters) {
  // one warp: wait until every owner is done with the buffers of scan seq - kShardDepth (same slot),
  // decide this rank's slice, re-arm the front counters (the tables are re-armed by
  // shard_alloc_kernel, which visits every bucket anyway)
  if (threadIdx.x < world && seq > kShardDepth) {
    while (ld_acquire_sys(&hdr->consumed[threadIdx.x]) + kShardDepth < seq) __nanosleep(64);
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // ── this rank's slice of the scan: every owner published, with consumed[d] = seq -
    //    kShardDepth, how many cells of its stripe that scan touched — the same numbers on every
    //    rank, so every rank derives the same split (shard_slice_plan, device_types.h).  Before
    //    any load is known (the first kShardDepth scans): all zero = equal slices. ──
    uint32_t load[kMaxShards];
    for (int d = 0; d < world; ++d)
      load[d] = seq > kShardDepth ? ld_relaxed_sys(&hdr->load[seq % kShardDepth][d]) : 0u;
    *slice_out = shard_slice_plan(load, world, rank, n_scan, back_weight);
  }
  if (blockIdx.x == 0 && threadIdx.x < kFrontCounterWords) counters[threadIdx.x] = 0u;
}

// FRONT: bucket segment allocation for every (stripe, bucket) of this rank's slice — the same
// two-atomics-per-warp scheme as commit_move_clear_kernel's group A
__global__ void __launch_bounds__(256)
shard_alloc_kernel(const __grid_constant__ TileBuffers tb, uint32_t bps, uint32_t* __restrict__ counters) {
  // blockIdx.y = the owner whose buckets this CTA row places: slots are handed out per owner,
  // because a bucket's records go into this source's area of ITS owner's arena
  const int lane = threadIdx.x & 31;
  const uint32_t owner = blockIdx.y;
  const size_t g0 = static_cast<size_t>(owner) * bps;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t b = tid; b - lane < bps; b += nthreads) {
    const uint32_t cnt = b < bps ? tb.bucket_count[g0 + b] : 0u;
    if (b < bps) {
      // re-arm: the histogram for this rank's next scan, the record cursors (peer-visible;
      // the owners finished reading this parity's — shard_begin_kernel waited for that)
      if (cnt) tb.bucket_count[g0 + b] = 0u;
      tb.bucket_cursor[g0 + b] = 0u;
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t slot0 = 0;
    if (lane == 0 && total) slot0 = atomicAdd(&counters[kShardSlotBase + owner], total);
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (cnt) tb.bucket_offset[g0 + b] = slot0 + incl - cnt;
  }
}

// FRONT, last kernel: tell every owner that this rank's records of scan `seq` are in place (the
// kernel boundary behind the scatter has completed its stores into the owners' arenas — cheaper
// than a system-scope fence in each of its thousands of CTAs, which was measured: +14 us; the
// flag is a system-scope release)
__global__ void shard_publish_front_kernel(const __grid_constant__ ShardFrontArgs a,
                                           const uint32_t* __restrict__ counters) {
  const int d = threadIdx.x;
  if (d >= a.world) return;
  const uint32_t inside = counters[CNT_INSIDE];
  a.peer_hdr[d]->inside[a.seq % kShardDepth][a.rank] = inside;
  __threadfence_system();
  st_release_sys(&a.peer_hdr[d]->ready[a.rank], a.seq);
}

// BACK, first kernel on every rank: wait for every source's front half, then (a) list the
// non-empty buckets of this rank's stripe with their record pieces, (b) do the map-side
// bookkeeping the one-GPU pipeline does in K2 / scatter: the obstacle reset of the last observing
// scan's cells, the touched-list hand-over and the sticky state flags — all decided on the
// scan-wide observation count, as the reference's single map would.
__global__ void __launch_bounds__(256)
shard_gather_kernel(const __grid_constant__ ShardBackArgs a, uint32_t* __restrict__ counters) {
  __shared__ uint32_t s_inside;
  if (threadIdx.x < a.world)
    while (ld_acquire_sys(&a.hdr->ready[threadIdx.x]) < a.seq) __nanosleep(64);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int s = 0; s < a.world; ++s) t += ld_relaxed_sys(&a.hdr->inside[a.seq % kShardDepth][s]);
    s_inside = t;
  }
  __syncthreads();
  const uint32_t scan_inside = s_inside;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  const uint32_t prev = a.st_cur->touched_count;
  const uint32_t flags_in = a.st_cur->flags;
  if (tid == 0) {
    // the scan's statistics as this rank saw them: kept / inside points of its slice
    counters[CNT_KEPT] = a.front_counters[CNT_KEPT];
    counters[CNT_INSIDE] = a.front_counters[CNT_INSIDE];
    a.st_out->geom = a.st_cur->geom;  // GLOBAL maps never move
    a.st_out->touched_count = scan_inside > 0 ? 0u : prev;
    uint32_t flags = flags_in;
    if (scan_inside > 0) flags = (flags | a.flags_if_cells) & ~static_cast<uint32_t>(SF_OBSTACLE_DIRTY);
    a.st_out->flags = flags;
  }
  if (scan_inside > 0 && a.obstacle) {
    if (flags_in & SF_OBSTACLE_DIRTY) {
      for (size_t c = tid; c < a.obstacle_cells; c += nthreads) a.obstacle[c] = nan_f32();
    } else {
      for (size_t j = tid; j < prev; j += nthreads) {
        const uint32_t k = a.touched_keys[j];
        if (k != a.invalid_key) a.obstacle[k] = nan_f32();
      }
    }
  }
  // job list: one thread per bucket of the stripe; the sources' counts and first slots (final
  // before the ready flag was released) are `2 * world` independent remote 4-byte reads, all in
  // flight together; one list-slot atomic per warp
  const uint32_t g0 = static_cast<uint32_t>(a.rank) * a.bps;
  const int lane = threadIdx.x & 31;
  for (size_t b0 = tid - lane; b0 < a.bps; b0 += nthreads) {
    const size_t b = b0 + lane;
    ShardJob J;
    J.bucket = static_cast<uint32_t>(b);
    J.total = 0;
#pragma unroll
    for (int s = 0; s < kMaxShards; ++s) {
      const bool on = s < a.world && b < a.bps;
      J.cnt[s] = on ? ld_relaxed_sys(&a.peer_cursor[s][g0 + b]) : 0u;
      J.off[s] = on ? ld_relaxed_sys(&a.peer_offset[s][g0 + b]) : 0u;
    }
#pragma unroll
    for (int s = 0; s < kMaxShards; ++s) J.total += J.cnt[s];
    const uint32_t live = __ballot_sync(0xffffffffu, J.total != 0u);
    if (!live) continue;
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(&counters[CNT_BUCKETS], static_cast<uint32_t>(__popc(live)));
    slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(live & ((1u << lane) - 1u));
    if (J.total) a.jobs[slot] = J;
  }
}

template <int BITS>
uint32_t tile_grid(uint32_t n_buckets) {
  const uint32_t cap = 148u * static_cast<uint32_t>(TileCfg<BITS>::kMinBlocks);
  return n_buckets < cap ? n_buckets : cap;
}

}  // namespace

#ifdef FDEM_PROBES
int tile_estimate_debug_cta_ns(unsigned long long* out1024) {
  return static_cast<int>(cudaMemcpyFromSymbol(out1024, g_k3t_cta_ns, sizeof(unsigned long long) * 1024));
}

int tile_estimate_debug_clocks(long long* out16) {
  return static_cast<int>(cudaMemcpyFromSymbol(out16, g_k3t_clocks, sizeof(long long) * 16));
}
#endif

int tile_estimate_configure() {
  cudaError_t e = cudaFuncSetAttribute(tile_estimate_kernel<8>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(TileSmem<8>)));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(tile_estimate_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(sizeof(TileSmem<9>)));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(tile_estimate_shard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(sizeof(TileSmem<10>)));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(tile_estimate_light_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(kLightSmem));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(tile_estimate_light_shard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(kLightSmem));
  if (e != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(cudaFuncSetAttribute(tile_estimate_kernel<10>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(sizeof(TileSmem<10>))));
}

void launch_scatter_records(const ScatterParams& p, cudaStream_t s, LaunchCounter& lc) {
  if (p.n == 0) return;
  scatter_records_kernel<<<(p.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(p);
  ++lc.mine;
}

void launch_shard_begin(ShardHeader* hdr, uint32_t seq, int world, int rank, uint32_t n_scan,
                        uint32_t back_weight, ShardSlice* slice_out, uint32_t* counters, cudaStream_t s,
                        LaunchCounter& lc) {
  shard_begin_kernel<<<1, 32, 0, s>>>(hdr, seq, world, rank, n_scan, back_weight, slice_out, counters);
  ++lc.mine;
}
void launch_shard_publish_front(const ShardFrontArgs& a, const uint32_t* counters, cudaStream_t s,
                                LaunchCounter& lc) {
  shard_publish_front_kernel<<<1, 32, 0, s>>>(a, counters);
  ++lc.mine;
}
void launch_shard_alloc(const TileBuffers& tb, uint32_t bps, int world, uint32_t* counters, cudaStream_t s,
                        LaunchCounter& lc) {
  shard_alloc_kernel<<<dim3(148 / world + 1, world), 256, 0, s>>>(tb, bps, counters);
  ++lc.mine;
}
void launch_shard_gather(const ShardBackArgs& a, uint32_t* counters, cudaStream_t s, LaunchCounter& lc) {
  shard_gather_kernel<<<148, 256, 0, s>>>(a, counters);
  ++lc.mine;
}
void launch_tile_estimate_light(const EstimateParams& p, const TileBuffers& tb, uint32_t* counters,
                                DeviceState* st_out, cudaStream_t s, LaunchCounter& lc) {
  tile_estimate_light_kernel<<<kLightGrid, kLightWarps * 32, kLightSmem, s>>>(p, tb, counters, st_out);
  ++lc.mine;
}
void launch_tile_estimate_light_shard(const EstimateParams& p, const ShardBackArgs& a, uint32_t* counters,
                                      DeviceState* st_out, cudaStream_t s, LaunchCounter& lc) {
  tile_estimate_light_shard_kernel<<<kLightGrid, kLightWarps * 32, kLightSmem, s>>>(p, a, counters, st_out);
  ++lc.mine;
}
void launch_tile_estimate_shard(const EstimateParams& p, const ShardBackArgs& a, uint32_t* counters,
                                DeviceState* st_out, const PublishArgs& pub, cudaStream_t s,
                                LaunchCounter& lc) {
  const uint32_t cap = 148u * static_cast<uint32_t>(TileCfg<10>::kMinBlocks);
  tile_estimate_shard_kernel<<<a.bps < cap ? a.bps : cap, TileCfg<10>::kThr, sizeof(TileSmem<10>), s>>>(
      p, a, counters, st_out, pub);
  ++lc.mine;
}

void launch_tile_estimate(const EstimateParams& p, const TileBuffers& tb, uint32_t* counters,
                          DeviceState* st_out, const PublishArgs& pub, cudaStream_t s,
                          LaunchCounter& lc) {
  // CTAs stride over the non-empty-bucket list; one wave of co-resident CTAs at most
  if (tb.bucket_bits == 8) {
    tile_estimate_kernel<8><<<tile_grid<8>(tb.n_buckets), TileCfg<8>::kThr, sizeof(TileSmem<8>), s>>>(
        p, tb, counters, st_out, pub);
  } else if (tb.bucket_bits == 9) {
    tile_estimate_kernel<9><<<tile_grid<9>(tb.n_buckets), TileCfg<9>::kThr, sizeof(TileSmem<9>), s>>>(
        p, tb, counters, st_out, pub);
  } else {
    tile_estimate_kernel<10><<<tile_grid<10>(tb.n_buckets), TileCfg<10>::kThr, sizeof(TileSmem<10>), s>>>(
        p, tb, counters, st_out, pub);
  }
  ++lc.mine;
}

}  // namespace fdem

namespace fdem {
KernelDesc desc_scatter_records(uint32_t n) {
  return KernelDesc{reinterpret_cast<const void*>(&scatter_records_kernel),
                    dim3((n + kThreads - 1) / kThreads), dim3(kThreads), 0};
}
KernelDesc desc_shard_begin() {
  return KernelDesc{reinterpret_cast<const void*>(&shard_begin_kernel), dim3(1), dim3(32), 0};
}
KernelDesc desc_shard_publish_front() {
  return KernelDesc{reinterpret_cast<const void*>(&shard_publish_front_kernel), dim3(1), dim3(32), 0};
}
KernelDesc desc_shard_alloc(int world) {
  return KernelDesc{reinterpret_cast<const void*>(&shard_alloc_kernel), dim3(148 / world + 1, world), dim3(256), 0};
}
KernelDesc desc_shard_gather() {
  return KernelDesc{reinterpret_cast<const void*>(&shard_gather_kernel), dim3(148), dim3(256), 0};
}
KernelDesc desc_tile_estimate_shard(uint32_t bps) {
  const uint32_t cap = 148u * static_cast<uint32_t>(TileCfg<10>::kMinBlocks);
  return KernelDesc{reinterpret_cast<const void*>(&tile_estimate_shard_kernel), dim3(bps < cap ? bps : cap),
                    dim3(TileCfg<10>::kThr), sizeof(TileSmem<10>)};
}
KernelDesc desc_tile_estimate_light() {
  return KernelDesc{reinterpret_cast<const void*>(&tile_estimate_light_kernel), dim3(kLightGrid),
                    dim3(kLightWarps * 32), kLightSmem};
}
KernelDesc desc_tile_estimate_light_shard() {
  return KernelDesc{reinterpret_cast<const void*>(&tile_estimate_light_shard_kernel), dim3(kLightGrid),
                    dim3(kLightWarps * 32), kLightSmem};
}
KernelDesc desc_tile_estimate(uint32_t n_buckets, uint32_t bucket_bits) {
  if (bucket_bits == 8)
    return KernelDesc{reinterpret_cast<const void*>(&tile_estimate_kernel<8>),
                      dim3(tile_grid<8>(n_buckets)), dim3(TileCfg<8>::kThr), sizeof(TileSmem<8>)};
  if (bucket_bits == 9)
    return KernelDesc{reinterpret_cast<const void*>(&tile_estimate_kernel<9>),
                      dim3(tile_grid<9>(n_buckets)), dim3(TileCfg<9>::kThr), sizeof(TileSmem<9>)};
  return KernelDesc{reinterpret_cast<const void*>(&tile_estimate_kernel<10>),
                    dim3(tile_grid<10>(n_buckets)), dim3(TileCfg<10>::kThr), sizeof(TileSmem<10>)};
}
}  // namespace fdem
