// Device side of the matcher.  Three kernels per (sub-)batch of image pairs:
//
//   score_*      N1 x N2 exact u8*u8->s32 dot products per pair, never written to memory: each
//                128 x 256 accumulator tile is scanned in place and only entries >= min_score (the
//                integer pre-filter derived from the acos table, see smb.cu derive_filter) are fed to
//   top2_insert  order-independent best / second-best accumulators per row AND per column
//                (64-bit keys: score << 32 | ~index, so max == "highest score, lowest index":
//                COLMAP's strict '>' ascending scan; the runner-up key's score is the second-best of
//                the multiset).
//   decide       per pair: acos-table distance test, ratio test ('>='), cross-check, ordered
//                compaction into FeatureMatch {idx1, idx2} rows (ascending idx1).
//
// Reference semantics: COLMAP 3.5 feature/sift.cc ComputeSiftDistanceMatrix /
// FindBestMatchesOneWay / FindBestMatches, called at
// /root/reference/integration/op_cpp/sequential_matching.cc:154.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace smb {

constexpr int kDim = 128;            // descriptor bytes == GEMM K
constexpr int kStripRows = 256;      // rows of image 1 per work item (two UMMA M=128 row blocks)
constexpr int kTileCols = 256;       // columns (rows of image 2) per accumulator tile == UMMA N
constexpr int kRowPad = 256;         // every cached image occupies a multiple of this many pool rows
constexpr int kLutSize = 512 * 512 + 1;

struct PairMeta {
  uint32_t a_row0;   // pool row of image 1
  uint32_t n1;
  uint32_t b_row0;   // pool row of image 2
  uint32_t n2;
  uint32_t acc_off;  // accumulator slots: rows at [acc_off, acc_off+n1), columns at [acc_off+n1, acc_off+n1+n2)
  uint32_t out_slot; // index into the per-call pair_out array
};

struct WorkItem {     // one strip (<= 256 rows) of image 1 against all of image 2
  uint32_t a_row;     // pool row of the strip
  uint32_t b_row;     // pool row of image 2
  uint32_t n_btiles;  // number of 256-column tiles
  uint32_t pair;      // index into PairMeta (batch-local)
  uint32_t m_tiles;   // 1 or 2 accumulator row blocks (128 rows each) in this strip
};

struct TopTwo {
  unsigned long long k1;  // best key
  unsigned long long k2;  // runner-up key
};

struct PairOut {
  uint32_t start;  // offset into the match buffer
  uint32_t count;
};

// key = score << 32 | ~slot.  "slot" is the accumulator slot of the OTHER axis (column slot in a row
// accumulator and vice versa); inside one pair slots are ordered like indices, so max(key) is "highest
// score, then lowest index" and decide() recovers the index by subtracting the pair's base slot.
__device__ __forceinline__ unsigned long long make_key(uint32_t score, uint32_t slot) {
  return (static_cast<unsigned long long>(score) << 32) | static_cast<uint32_t>(~slot);
}

// Insert into {best, runner-up} accumulators of a row and a column with atomic max.  Keys are
// distinct (the slot is part of the key); k1 ends as the maximum, and every insertion hands
// min(previous best, key) to k2, so k2 ends as the second-largest key whatever the order.
__device__ __forceinline__ void top2_insert2(TopTwo* __restrict__ acc, uint32_t row_slot, uint32_t col_slot,
                                             uint32_t score) {
  TopTwo* tr = acc + row_slot;
  TopTwo* tc = acc + col_slot;
  const unsigned long long kr = make_key(score, col_slot), kc = make_key(score, row_slot);
  const unsigned long long o_r = atomicMax(&tr->k1, kr);  // both round trips in flight together
  const unsigned long long o_c = atomicMax(&tc->k1, kc);
  const unsigned long long l_r = o_r < kr ? o_r : kr;
  const unsigned long long l_c = o_c < kc ? o_c : kc;
  if (l_r) atomicMax(&tr->k2, l_r);
  if (l_c) atomicMax(&tc->k2, l_c);
}

// =====================================================================================
// Production score kernel: TMA -> swizzled smem -> tcgen05.mma kind::i8 -> TMEM -> filter -> rescore
//
//   warp 0        TMA producer: A strip (256 rows, kept for the whole item) + ring of B tiles (256 rows)
//   warp 1        MMA issuer: per B tile two accumulator tiles (strip rows 0-127 / 128-255), 4 x K32 each,
//                 ping-pong between the two 256-column halves of TMEM
//   warps 4-11    filter epilogue: tcgen05.ld + 3-input max tree over each thread's 32-column run; the
//                 accumulator buffer is released as soon as it has been read.  A run whose maximum
//                 reaches min_score is not re-read: its (row, 32-column run) address goes to a
//                 shared-memory queue.
//   warps 2,3,12,13  rescoring: pop up to 4 (row, run) entries at a time, recompute their 32 dot products
//                 exactly on CUDA cores (__dp4a over the descriptors in L2) and feed every score
//                 >= min_score to the top-2 accumulators (global atomics).  All of that latency is off
//                 the tile pipeline.
// =====================================================================================
constexpr int kStages = 3;                     // B-tile ring
constexpr int kAStages = 2;                    // A-strip ring (next item's strip prefetched)
constexpr int kMTile = 128;                    // UMMA M
constexpr int kABytes = kStripRows * kDim;     // 32 KiB (two 128-row boxes)
constexpr int kBBytes = kTileCols * kDim;      // 32 KiB (two 128-row boxes)
constexpr int kEpiWarps = 8;                   // 2 per TMEM lane quarter
constexpr int kRescoreWarps = 4;               // warps 2, 3, 12, 13
constexpr int kScoreWarps = 4 + kEpiWarps + 2;
constexpr int kScoreThreads = 32 * kScoreWarps;
constexpr int kRunCols = 32;                   // columns per thread per tcgen05.ld == rescoring granularity
constexpr int kHitSlots = 12;                  // staging ring for (row, run) hits
constexpr int kHitBytes = kDim + kRunCols * kDim;  // one image-1 descriptor + 32 image-2 descriptors

struct ScoreShared {
  uint64_t a_full[kAStages], a_empty[kAStages];
  uint64_t b_full[kStages], b_empty[kStages];
  uint64_t t_full[2], t_empty[2];
  uint64_t h_full[kHitSlots];      // hit staging: descriptors of the hit have landed (tx-count barrier)
  uint32_t h_free_gen[kHitSlots];  // generation that may next write the slot
  uint2 h_meta[kHitSlots];         // {accumulator slot of the row, accumulator slot of the run's first column}
  uint32_t h_tail;                 // next hit sequence number
  uint32_t h_done;                 // epilogue warps that have finished
  uint32_t tmem_base;
  uint32_t pad_;
};
constexpr int kScoreSmemBytes =
    1024 /*align slack*/ + kAStages * kABytes + kStages * kBBytes + kHitSlots * kHitBytes + (int)sizeof(ScoreShared);
static_assert(kScoreSmemBytes <= 227 * 1024, "shared memory budget");

// max over 32 accumulator entries with 3-input integer max (VIMNMX3), as a tree for ILP:
// 16 instructions per 32 entries
__device__ __forceinline__ int max_tree32(const uint32_t (&v)[32]) {
  int a[11];
#pragma unroll
  for (int g = 0; g < 10; ++g) a[g] = __vimax3_s32((int)v[3 * g], (int)v[3 * g + 1], (int)v[3 * g + 2]);
  a[10] = max((int)v[30], (int)v[31]);
  const int b0 = __vimax3_s32(a[0], a[1], a[2]);
  const int b1 = __vimax3_s32(a[3], a[4], a[5]);
  const int b2 = __vimax3_s32(a[6], a[7], a[8]);
  const int b3 = max(a[9], a[10]);
  return max(__vimax3_s32(b0, b1, b2), b3);
}

__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// Epilogue side: a (row, 32-column run) whose maximum reached min_score.  Claim a staging slot and let
// the bulk-copy engine fetch the 1 + 32 descriptors involved (L2 -> shared memory); the rescoring warps
// pick the slot up when its barrier completes.  Nothing here waits on memory.
__device__ __forceinline__ void hit_push(ScoreShared* sh, uint32_t smem_hits, const uint8_t* __restrict__ pool,
                                         uint32_t a_row, uint32_t b_row, uint32_t row_slot, uint32_t col_slot0) {
  const uint32_t s = atomicAdd(&sh->h_tail, 1u);
  const uint32_t slot = s % kHitSlots, gen = s / kHitSlots;
  volatile uint32_t* free_gen = &sh->h_free_gen[slot];
  uint32_t spins = 0;
  while (*free_gen != gen) {  // ring full: back-pressure until the rescoring warps catch up
    if (++spins > SMB_MBAR_SPIN_LIMIT) __trap();
  }
  sh->h_meta[slot] = make_uint2(row_slot, col_slot0);
  const uint32_t bar = ptx::smem_u32(&sh->h_full[slot]);
  const uint32_t dst = smem_hits + slot * kHitBytes;
  ptx::mbar_arrive_expect_tx(bar, kHitBytes);  // release: the meta store is visible to whoever sees the phase flip
  ptx::bulk_load(dst, pool + (size_t)a_row * kDim, kDim, bar);
  ptx::bulk_load(dst + kDim, pool + (size_t)b_row * kDim, kRunCols * kDim, bar);
}

// Rescoring warp r owns hit sequence numbers r, r + kRescoreWarps, ...  Lane l recomputes the exact score
// of column (run + l) from the staged descriptors (bank-conflict-free rotation); scores >= min_score go to
// the top-2 accumulators.
__device__ __forceinline__ void rescore_loop(ScoreShared* sh, uint32_t smem_hits, TopTwo* __restrict__ acc, int min_score,
                                             uint32_t r, uint32_t lane, unsigned long long* cand_counter) {
  uint32_t n = r, count = 0;
  volatile uint32_t* tail = &sh->h_tail;
  volatile uint32_t* done = &sh->h_done;
  for (;;) {
    const uint32_t slot = n % kHitSlots, gen = n / kHitSlots;
    const uint32_t bar = ptx::smem_u32(&sh->h_full[slot]);
    bool finished = false;
    while (!ptx::mbar_try_wait(bar, gen & 1)) {
      if (*done == kEpiWarps) {
        fence_cta();
        if (n >= *tail) {  // h_done is bumped only after every push of that warp is visible
          finished = true;
          break;
        }
      }
      __nanosleep(128);
    }
    if (finished) break;
    uint2 meta;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];"
                 : "=r"(meta.x), "=r"(meta.y)
                 : "r"(ptx::smem_u32(&sh->h_meta[slot]))
                 : "memory");
    const uint32_t base = smem_hits + slot * kHitBytes;
    uint32_t sc = 0;
#pragma unroll
    for (int k = 0; k < kDim / 4; ++k) {
      const uint32_t w = (lane + k) & 31;  // rotation: every lane touches a different bank in each step
      uint32_t av, bv;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(av) : "r"(base + w * 4));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(bv) : "r"(base + kDim + lane * kDim + w * 4));
      sc = __dp4a(av, bv, sc);
    }
    __syncwarp();
    if (lane == 0) sh->h_free_gen[slot] = gen + 1;  // slot may be refilled
    if ((int)sc >= min_score) {  // pool padding rows are zero and can never get here
      top2_insert2(acc, meta.x, meta.y + lane, sc);
      ++count;
    }
    n += kRescoreWarps;
  }
  if (cand_counter && count) atomicAdd(cand_counter, (unsigned long long)count);
}

// Bring-up instrumentation (dbg & 32): per-CTA cycle totals, [cta][16]:
// 0 mma: wait b_full, 1 mma: wait t_empty, 2 mma: total, 3 mma: tiles,
// 4 epi(warp 4): wait t_full, 5 epi: ld+max, 6 epi: push, 7 epi: total, 8 prod: wait b_empty, 9 prod: total
__device__ long long g_score_clocks[148 * 16];

__global__ void __launch_bounds__(kScoreThreads, 1)
score_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ pool,
                     const WorkItem* __restrict__ items, uint32_t n_items, const PairMeta* __restrict__ pairs,
                     TopTwo* __restrict__ acc, int min_score, unsigned long long* cand_counter, uint32_t dbg) {
  // dbg (bring-up timing experiments only, results become meaningless): 1 = epilogue releases tiles unread,
  // 2 = B tiles are not loaded, 4 = hits are not queued
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024 B alignment
  const uint32_t smem_a = smem0;
  const uint32_t smem_b = smem0 + kAStages * kABytes;
  const uint32_t smem_hits = smem_b + kStages * kBBytes;
  ScoreShared* sh = reinterpret_cast<ScoreShared*>(smem_raw + (smem0 - ptx::smem_u32(smem_raw)) + kAStages * kABytes +
                                                   kStages * kBBytes + kHitSlots * kHitBytes);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap);
    for (int s = 0; s < kAStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sh->a_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&sh->a_empty[s]), 1);
    }
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sh->b_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&sh->b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sh->t_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&sh->t_empty[s]), kEpiWarps);
    }
    for (int s = 0; s < kHitSlots; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sh->h_full[s]), 1);
      sh->h_free_gen[s] = 0;
    }
    sh->h_tail = 0;
    sh->h_done = 0;
    ptx::fence_barrier_init();
  }
  if (warp == 2) {  // whole warp: TMEM allocation (all 512 columns: two 256-column accumulators)
    ptx::tmem_alloc_512(ptx::smem_u32(&sh->tmem_base));
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      long long c_wait = 0, c_t0 = clock64();
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = items[it];
        ptx::mbar_wait(ptx::smem_u32(&sh->a_empty[as]), aph ^ 1);
        const uint32_t afull = ptx::smem_u32(&sh->a_full[as]);
        ptx::mbar_arrive_expect_tx(afull, w.m_tiles * (kABytes / 2));
        for (uint32_t mh = 0; mh < w.m_tiles; ++mh)
          ptx::tma_load_2d(smem_a + as * kABytes + mh * (kABytes / 2), &tmap, afull, 0, (int32_t)(w.a_row + mh * kMTile));
        if (++as == kAStages) { as = 0; aph ^= 1; }
        for (uint32_t t = 0; t < w.n_btiles; ++t) {
          const long long c0 = (dbg & 32) ? clock64() : 0;
          ptx::mbar_wait(ptx::smem_u32(&sh->b_empty[bs]), bph ^ 1);
          if (dbg & 32) c_wait += clock64() - c0;
          const uint32_t full = ptx::smem_u32(&sh->b_full[bs]);
          if (dbg & 2) {
            ptx::mbar_arrive(full);
          } else {
            ptx::mbar_arrive_expect_tx(full, kBBytes);
            const int32_t r = (int32_t)(w.b_row + t * kTileCols);
            ptx::tma_load_2d(smem_b + bs * kBBytes, &tmap, full, 0, r);
            ptx::tma_load_2d(smem_b + bs * kBBytes + kBBytes / 2, &tmap, full, 0, r + kTileCols / 2);
          }
          if (++bs == kStages) { bs = 0; bph ^= 1; }
        }
      }
      if ((dbg & 32) && blockIdx.x < 148) {
        g_score_clocks[blockIdx.x * 16 + 8] = c_wait;
        g_score_clocks[blockIdx.x * 16 + 9] = clock64() - c_t0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_u8u8s32(kMTile, kTileCols);
      uint32_t as = 0, aph = 0, bs = 0, bph = 0, ts = 0, tph = 0;
      long long c_b = 0, c_t = 0, c_n = 0, c_t0 = clock64();
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const uint32_t n_btiles = items[it].n_btiles, m_tiles = items[it].m_tiles;
        ptx::mbar_wait(ptx::smem_u32(&sh->a_full[as]), aph);
        const uint64_t adesc0 = ptx::make_kmajor_sw128_desc(smem_a + as * kABytes);
        for (uint32_t t = 0; t < n_btiles; ++t) {
          long long c0 = (dbg & 32) ? clock64() : 0;
          ptx::mbar_wait(ptx::smem_u32(&sh->b_full[bs]), bph);
          if (dbg & 32) c_b += clock64() - c0;
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(smem_b + bs * kBBytes);
          for (uint32_t mh = 0; mh < m_tiles; ++mh) {
            c0 = (dbg & 32) ? clock64() : 0;
            ptx::mbar_wait(ptx::smem_u32(&sh->t_empty[ts]), tph ^ 1);
            if (dbg & 32) { c_t += clock64() - c0; ++c_n; }
            ptx::tcgen05_fence_after();
            const uint64_t adesc = adesc0 + mh * ((kABytes / 2) >> 4);
            const uint32_t d = tmem_base + ts * kTileCols;
#pragma unroll
            for (uint32_t k = 0; k < kDim / 32; ++k)  // UMMA K = 32 bytes; advance inside the swizzle atom
              ptx::umma_i8(d, adesc + k * 2, bdesc + k * 2, idesc, k);
            ptx::umma_commit(ptx::smem_u32(&sh->t_full[ts]));
            if (++ts == 2) { ts = 0; tph ^= 1; }
          }
          ptx::umma_commit(ptx::smem_u32(&sh->b_empty[bs]));
          if (++bs == kStages) { bs = 0; bph ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&sh->a_empty[as]));
        if (++as == kAStages) { as = 0; aph ^= 1; }
      }
      if ((dbg & 32) && blockIdx.x < 148) {
        g_score_clocks[blockIdx.x * 16 + 0] = c_b;
        g_score_clocks[blockIdx.x * 16 + 1] = c_t;
        g_score_clocks[blockIdx.x * 16 + 2] = clock64() - c_t0;
        g_score_clocks[blockIdx.x * 16 + 3] = c_n;
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ------------------------------------------------------------ filter epilogue (8 warps)
    const uint32_t quarter = warp & 3;            // TMEM lanes [32*quarter, +32) are visible to this warp
    const uint32_t half = (warp - 4) >> 2;        // which 128 of the tile's 256 columns
    const uint32_t lane_addr = (quarter * 32u) << 16;
    uint32_t ts = 0, tph = 0;
    long long c_w = 0, c_l = 0, c_p = 0, c_t0 = clock64();
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const WorkItem w = items[it];
      const PairMeta pm = pairs[w.pair];
      const uint32_t a_row = w.a_row + quarter * 32 + lane;              // pool row of this thread's descriptor (mh = 0)
      const uint32_t row_slot = pm.acc_off + (a_row - pm.a_row0);        // its accumulator slot
      const uint32_t col_slot0 = pm.acc_off + pm.n1 + half * 128u;
      for (uint32_t t = 0; t < w.n_btiles; ++t) {
        for (uint32_t mh = 0; mh < w.m_tiles; ++mh) {
          long long c0 = (dbg & 32) ? clock64() : 0;
          ptx::mbar_wait(ptx::smem_u32(&sh->t_full[ts]), tph);
          long long c1 = (dbg & 32) ? clock64() : 0;
          ptx::tcgen05_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + ts * kTileCols + half * 128u;
          int mc[4] = {0, 0, 0, 0};
          if (!(dbg & 1)) {
            uint32_t v0[32], v1[32], v2[32], v3[32];  // all four runs in flight; ptxas tracks each load's registers
            ptx::tmem_ld_32x32b_x32(taddr, v0);
            ptx::tmem_ld_32x32b_x32(taddr + kRunCols, v1);
            ptx::tmem_ld_32x32b_x32(taddr + 2 * kRunCols, v2);
            ptx::tmem_ld_32x32b_x32(taddr + 3 * kRunCols, v3);
            ptx::tmem_wait_ld();
            mc[0] = max_tree32(v0);
            mc[1] = max_tree32(v1);
            mc[2] = max_tree32(v2);
            mc[3] = max_tree32(v3);
          }
          // the accumulator values are in registers: hand the TMEM buffer back to the MMA warp
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&sh->t_empty[ts]));
          if (++ts == 2) { ts = 0; tph ^= 1; }
          const int m = max(max(mc[0], mc[1]), max(mc[2], mc[3]));
          long long c2 = (dbg & 32) ? clock64() + (m & 0) : 0;
          if (__any_sync(0xffffffffu, m >= min_score) && !(dbg & 4)) {
            const uint32_t j0 = t * kTileCols + half * 128u;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (mc[c] >= min_score)
                hit_push(sh, smem_hits, pool, a_row + mh * kMTile, w.b_row + j0 + c * kRunCols, row_slot + mh * kMTile,
                         col_slot0 + t * kTileCols + c * kRunCols);
          }
          if (dbg & 32) { c_w += c1 - c0; c_l += c2 - c1; c_p += clock64() - c2; }
        }
      }
    }
    __syncwarp();
    if ((dbg & 32) && warp == 4 && lane == 0 && blockIdx.x < 148) {
      g_score_clocks[blockIdx.x * 16 + 4] = c_w;
      g_score_clocks[blockIdx.x * 16 + 5] = c_l;
      g_score_clocks[blockIdx.x * 16 + 6] = c_p;
      g_score_clocks[blockIdx.x * 16 + 7] = clock64() - c_t0;
    }
    if (lane == 0) {
      fence_cta();
      atomicAdd(&sh->h_done, 1u);
    }
  } else {
    // ------------------------------------------------------------ rescoring (warps 2, 3, 12, 13)
    rescore_loop(sh, smem_hits, acc, min_score, warp < 4 ? warp - 2 : warp - 10, lane, cand_counter);
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_512(tmem_base);
  }
}

// =====================================================================================
// Test-only device cross-check: the same contract on CUDA cores (__dp4a), no tensor cores,
// no TMA.  Selected with SMB_ENGINE_DP4A; never the default.
// =====================================================================================
constexpr int kDp4aThreads = 256;
constexpr int kDp4aCols = 64;

__global__ void __launch_bounds__(kDp4aThreads)
score_dp4a_kernel(const uint8_t* __restrict__ pool, const WorkItem* __restrict__ items, uint32_t n_items,
                  const PairMeta* __restrict__ pairs, TopTwo* __restrict__ acc, int min_score,
                  unsigned long long* cand_counter) {
  __shared__ uint32_t sa[kMTile][kDim / 4 + 1];
  __shared__ uint32_t sb[kDp4aCols][kDim / 4 + 1];
  const uint32_t tid = threadIdx.x;
  const uint32_t ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, 8 rows x 4 columns each
  unsigned long long count = 0;
  for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const WorkItem w = items[it];
    const PairMeta pm = pairs[w.pair];
    for (uint32_t mh = 0; mh < w.m_tiles; ++mh) {
      const uint32_t a_row = w.a_row + mh * kMTile;
      const uint32_t* ga = reinterpret_cast<const uint32_t*>(pool + (size_t)a_row * kDim);
      __syncthreads();
      for (uint32_t x = tid; x < kMTile * (kDim / 4); x += kDp4aThreads) sa[x >> 5][x & 31] = ga[x];
      const uint32_t n_cols = w.n_btiles * kTileCols;
      for (uint32_t c0 = 0; c0 < n_cols; c0 += kDp4aCols) {
        const uint32_t* gb = reinterpret_cast<const uint32_t*>(pool + (size_t)(w.b_row + c0) * kDim);
        __syncthreads();
        for (uint32_t x = tid; x < kDp4aCols * (kDim / 4); x += kDp4aThreads) sb[x >> 5][x & 31] = gb[x];
        __syncthreads();
        uint32_t s[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) s[r][c] = 0;
        for (int k = 0; k < kDim / 4; ++k) {
          uint32_t a[8], b[4];
#pragma unroll
          for (int r = 0; r < 8; ++r) a[r] = sa[ty * 8 + r][k];
#pragma unroll
          for (int c = 0; c < 4; ++c) b[c] = sb[tx * 4 + c][k];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[r][c] = __dp4a(a[r], b[c], s[r][c]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t i = a_row - pm.a_row0 + ty * 8 + r, j = c0 + tx * 4 + c;
            if ((int)s[r][c] >= min_score && i < pm.n1 && j < pm.n2) {
              top2_insert2(acc, pm.acc_off + i, pm.acc_off + pm.n1 + j, s[r][c]);
              ++count;
            }
          }
      }
    }
  }
  if (cand_counter && count) atomicAdd(cand_counter, count);
}

// =====================================================================================
// decide: FindBestMatchesOneWay tests + cross-check + ordered compaction, one CTA per pair
// =====================================================================================
constexpr int kDecideThreads = 512;

// acosf(min(score / 512^2, 1)) through the host-libm table; rows/columns without any surviving
// score (k1 == 0) never match.  Returns the matched index or -1.
__device__ __forceinline__ int decide_one(const TopTwo t, uint32_t base_slot, const float* __restrict__ lut,
                                          float max_ratio, float max_distance) {
  const uint32_t best = static_cast<uint32_t>(t.k1 >> 32);
  if (best == 0) return -1;
  const uint32_t second = static_cast<uint32_t>(t.k2 >> 32);
  const float bn = __ldg(lut + min(best, (uint32_t)(kLutSize - 1)));
  if (bn > max_distance) return -1;
  const float sn = __ldg(lut + min(second, (uint32_t)(kLutSize - 1)));
  if (bn >= __fmul_rn(max_ratio, sn)) return -1;  // '>=' rejects best == second-best
  return static_cast<int>(~static_cast<uint32_t>(t.k1) - base_slot);
}

__global__ void __launch_bounds__(kDecideThreads)
decide_kernel(const PairMeta* __restrict__ pairs, TopTwo* __restrict__ acc, const float* __restrict__ lut,
              float max_ratio, float max_distance, int cross_check, uint2* __restrict__ out /* FeatureMatch */,
              unsigned long long* __restrict__ out_total, PairOut* __restrict__ pair_out) {
  __shared__ uint32_t warp_sums[kDecideThreads / 32];
  __shared__ uint32_t s_base;
  const PairMeta pm = pairs[blockIdx.x];
  TopTwo* rows = acc + pm.acc_off;
  TopTwo* cols = rows + pm.n1;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (cross_check) {
    for (uint32_t j = tid; j < pm.n2; j += kDecideThreads) {
      const int m21 = decide_one(cols[j], pm.acc_off, lut, max_ratio, max_distance);  // keys hold row slots
      cols[j].k1 = static_cast<unsigned long long>(static_cast<uint32_t>(m21));
    }
  }
  __syncthreads();

  // pass 1: decide every row, remember the verdict in place, count
  uint32_t cnt = 0;
  for (uint32_t i = tid; i < pm.n1; i += kDecideThreads) {
    int m12 = decide_one(rows[i], pm.acc_off + pm.n1, lut, max_ratio, max_distance);  // keys hold column slots
    if (m12 >= 0 && cross_check && static_cast<uint32_t>(cols[m12].k1) != i) m12 = -1;
    rows[i].k1 = static_cast<unsigned long long>(static_cast<uint32_t>(m12));
    cnt += (m12 >= 0);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) warp_sums[wid] = cnt;
  __syncthreads();
  if (tid == 0) {
    uint32_t total = 0;
    for (int w = 0; w < kDecideThreads / 32; ++w) total += warp_sums[w];
    const unsigned long long base = atomicAdd(out_total, (unsigned long long)total);
    s_base = static_cast<uint32_t>(base);
    pair_out[pm.out_slot].start = static_cast<uint32_t>(base);
    pair_out[pm.out_slot].count = total;
  }
  __syncthreads();
  uint32_t running = s_base;

  // pass 2: ordered write (ascending idx1), block-wide exclusive scan per chunk of rows
  for (uint32_t i0 = 0; i0 < pm.n1; i0 += kDecideThreads) {
    const uint32_t i = i0 + tid;
    int m12 = -1;
    if (i < pm.n1) m12 = static_cast<int>(static_cast<uint32_t>(rows[i].k1));
    const uint32_t ballot = __ballot_sync(0xffffffffu, m12 >= 0);
    __syncthreads();  // warp_sums reuse
    if (lane == 0) warp_sums[wid] = __popc(ballot);
    __syncthreads();
    uint32_t before = 0, chunk_total = 0;
#pragma unroll
    for (int w = 0; w < kDecideThreads / 32; ++w) {
      const uint32_t s = warp_sums[w];
      if (w < (int)wid) before += s;
      chunk_total += s;
    }
    if (m12 >= 0) {
      const uint32_t pos = running + before + __popc(ballot & ((1u << lane) - 1));
      out[pos] = make_uint2(i, static_cast<uint32_t>(m12));
    }
    running += chunk_total;
  }
}

}  // namespace smb
