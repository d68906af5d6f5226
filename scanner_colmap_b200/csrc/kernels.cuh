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
constexpr int kStripRows = 512;      // rows of image 1 per work item: a CTA pair (M = 256) x up to two row blocks
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

struct WorkItem {     // one strip (<= 512 rows) of image 1 against all of image 2, processed by one CTA pair
  uint32_t a_row;     // pool row of the strip
  uint32_t b_row;     // pool row of image 2
  uint32_t n_btiles;  // number of 256-column tiles
  uint32_t pair;      // index into PairMeta (batch-local)
  uint32_t m_tiles;   // 1 or 2 row blocks of 256 rows (128 per CTA of the pair) in this strip
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
// Production score kernel, one CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 512-row strip:
// TMA -> swizzled smem -> tcgen05.mma.cta_group::2 kind::i8 (M = 256 over the pair) -> TMEM -> filter -> rescore
//
// Why a pair: with one CTA the MMA reads A (4 KB) + B (8 KB) from shared memory every 128 cycles while
// TMA writes the next B tile, which saturates the 128 B/clk shared-memory port (measured: the tensor
// pipe stalls at ~60 %).  In a pair each CTA feeds its own 128 rows of A and only HALF of B (4 KB),
// and loads only half of every B tile from L2.
//
//   warp 0        TMA producer (both CTAs): own A rows (2 x 128, kept for the whole item) + ring of B
//                 half tiles (128 descriptors); byte counts are reported to the leader CTA's barriers
//   warp 1        MMA issuer (leader CTA only): per B tile two accumulator tiles (row blocks 0 / 1),
//                 4 x K32 each, ping-pong between the two 256-column halves of TMEM; completion is
//                 multicast to both CTAs' barriers
//   warps 4-11    filter epilogue (both CTAs, own 128 accumulator rows): tcgen05.ld + 3-input max tree
//                 over each thread's 32-column run; the accumulator buffer is released (to the leader)
//                 as soon as it has been read.  A run whose maximum reaches min_score is not re-read:
//                 its (row, 32-column run) address goes to a shared-memory staging ring.
//   warps 2,3     insert: feed survivors to the top-2 accumulators (global atomics, off the tile pipeline);
//                 runs with several survivors are first re-scored exactly on CUDA cores (__dp4a).
// =====================================================================================
constexpr int kStages = 6;                     // ring of B half tiles (128 descriptors each)
constexpr int kAStages = 2;                    // A ring (next item's rows prefetched)
constexpr int kMTile = 128;                    // accumulator rows per CTA (UMMA M = 256 over the pair)
constexpr int kABytes = 2 * kMTile * kDim;     // 32 KiB per CTA: its 128 rows of both row blocks
constexpr int kBBytes = (kTileCols / 2) * kDim;  // 16 KiB per CTA: its half of a 256-column B tile
constexpr int kEpiWarps = 16;                  // 4 per TMEM lane quarter, 64 accumulator columns each
constexpr int kEpiCols = kTileCols / (kEpiWarps / 4);
constexpr int kRescoreWarps = 2;               // insert / re-score warps
constexpr int kScoreWarps = 4 + kEpiWarps;     // 0 TMA, 1 MMA, 2-3 insert, 4-19 epilogue
constexpr int kScoreThreads = 32 * kScoreWarps;
constexpr int kRunCols = 32;                   // columns per thread per tcgen05.ld == rescoring granularity
constexpr int kRunsPerWarp = kEpiCols / kRunCols;
constexpr int kMailSlots = 32;                 // per epilogue warp: ring of pending records (>= 32: one post can carry 32)
constexpr int kStageSlots = 1;                 // per rescoring warp: descriptor staging buffer
constexpr int kHitBytes = kDim + kRunCols * kDim;  // one image-1 descriptor + 32 image-2 descriptors
constexpr uint32_t kDirectTag = 0xFFFFFFFFu;
static_assert(kRunsPerWarp == 2, "epilogue code is written for two 32-column runs per warp");

struct ScoreShared {
  uint64_t a_full[kAStages], a_empty[kAStages];
  uint64_t b_full[kStages], b_empty[kStages];
  uint64_t t_full[2], t_empty[2];
  uint64_t h_full[kRescoreWarps][kStageSlots];  // staging: the hit's descriptors have landed (tx-count barrier)
  uint32_t mail_head[kEpiWarps];   // records written by epilogue warp e (monotonic)
  uint32_t mail_tail[kEpiWarps];   // records consumed by its rescoring warp (monotonic)
  uint32_t epi_done;               // epilogue warps that have finished
  uint32_t tmem_base;
  // Records.  Survivor (the common case, exact score already known from the accumulator):
  //   {row accumulator slot, column accumulator slot, score, kDirectTag}
  // Run with several survivors (rare; re-scored on CUDA cores):
  //   {pool row of the image-1 descriptor, pool row of the run's first image-2 descriptor,
  //    row accumulator slot, accumulator slot of the run's first column}
  uint4 mail[kEpiWarps][kMailSlots];
};
constexpr int kScoreSmemBytes = 1024 /*align slack*/ + kAStages * kABytes + kStages * kBBytes +
                                kRescoreWarps * kStageSlots * kHitBytes + (int)sizeof(ScoreShared);
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
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(p)) : "memory");
  return v;
}

// Epilogue side (warp-uniform call): lanes with `hit` append one record each to this warp's mailbox.
// Single producer (this warp) / single consumer (its rescoring warp): no atomics, a few instructions.
__device__ __forceinline__ void mail_post(ScoreShared* sh, uint32_t e, uint32_t lane, bool hit, uint4 rec, uint32_t& head) {
  const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
  if (ballot == 0) return;
  if (hit) {
    const uint32_t pos = head + __popc(ballot & ((1u << lane) - 1));
    uint32_t spins = 0;
    while (pos - ld_volatile_shared(&sh->mail_tail[e]) >= (uint32_t)kMailSlots) {  // ring full: back-pressure
      if (++spins > (1u << 28)) __trap();
    }
    sh->mail[e][pos % kMailSlots] = rec;
  }
  head += __popc(ballot);
  fence_cta();
  __syncwarp();
  if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&sh->mail_head[e]) = head;
}

// Survivors of one 32-column run held in registers: bit e of the result is set iff v[e] >= min_score.
__device__ __forceinline__ uint32_t survivor_mask(const uint32_t (&v)[32], int min_score) {
  uint32_t mask = 0;
#pragma unroll
  for (int e = 0; e < 32; ++e) mask |= ((int)v[e] >= min_score) ? (1u << e) : 0u;
  return mask;
}

// Insert warp r serves the mailboxes of epilogue warps r, r + 2, r + 4, ...  Each lane takes one record.
// Survivor records go straight to the top-2 accumulators (32 insertions in flight per warp).  A record
// naming a run with several survivors is re-scored by the whole warp: the bulk-copy engine stages the
// 1 + 32 descriptors (L2 -> shared memory), lane l recomputes the exact score of column (run + l) with
// __dp4a (bank-conflict-free rotation) and inserts it if it reaches min_score.
__device__ __forceinline__ void insert_loop(ScoreShared* sh, uint32_t smem_stage, const uint8_t* __restrict__ pool,
                                            TopTwo* __restrict__ acc, int min_score, uint32_t r, uint32_t lane,
                                            unsigned long long* cand_counter, uint32_t dbg) {
  uint32_t tail[kEpiWarps / kRescoreWarps];
#pragma unroll
  for (int k = 0; k < kEpiWarps / kRescoreWarps; ++k) tail[k] = 0;
  uint32_t count = 0, staged = 0;
  const uint32_t bar = ptx::smem_u32(&sh->h_full[r][0]);
  const uint32_t base = smem_stage + r * kStageSlots * kHitBytes;
  for (;;) {
    bool any = false;
#pragma unroll
    for (int k = 0; k < kEpiWarps / kRescoreWarps; ++k) {
      const uint32_t e = r + k * kRescoreWarps;
      const uint32_t head = ld_volatile_shared(&sh->mail_head[e]);
      const uint32_t avail = head - tail[k];
      if (avail == 0) continue;
      any = true;
      fence_cta();
      const uint32_t n = avail < 32u ? avail : 32u;
      uint4 rec = make_uint4(0, 0, 0, 0);
      if (lane < n)
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rec.x), "=r"(rec.y), "=r"(rec.z), "=r"(rec.w)
                     : "r"(ptx::smem_u32(&sh->mail[e][(tail[k] + lane) % kMailSlots]))
                     : "memory");
      tail[k] += n;
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&sh->mail_tail[e]) = tail[k];  // entries may be reused
      const bool direct = lane < n && rec.w == kDirectTag;
      if (direct && !(dbg & 8)) {
        top2_insert2(acc, rec.x, rec.y, rec.z);
        ++count;
      }
      uint32_t multi = __ballot_sync(0xffffffffu, lane < n && rec.w != kDirectTag);
      while (multi) {  // rare: warp-cooperative re-scoring, one run at a time
        const int src = __ffs(multi) - 1;
        multi &= multi - 1;
        const uint32_t a_row = __shfl_sync(0xffffffffu, rec.x, src), b_row = __shfl_sync(0xffffffffu, rec.y, src);
        const uint32_t row_slot = __shfl_sync(0xffffffffu, rec.z, src), col_slot0 = __shfl_sync(0xffffffffu, rec.w, src);
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(bar, kHitBytes);
          ptx::bulk_load(base, pool + (size_t)a_row * kDim, kDim, bar);
          ptx::bulk_load(base + kDim, pool + (size_t)b_row * kDim, kRunCols * kDim, bar);
        }
        ptx::mbar_wait(bar, staged & 1);
        ++staged;
        uint32_t sc = 0;
#pragma unroll
        for (int q = 0; q < kDim / 4; ++q) {
          const uint32_t w = (lane + q) & 31;  // rotation: every lane touches a different bank in each step
          uint32_t av, bv;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(av) : "r"(base + w * 4));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(bv) : "r"(base + kDim + lane * kDim + w * 4));
          sc = __dp4a(av, bv, sc);
        }
        __syncwarp();  // every lane has read the staging buffer before it is refilled
        if ((int)sc >= min_score) {  // pool padding rows are zero and can never get here
          top2_insert2(acc, row_slot, col_slot0 + lane, sc);
          ++count;
        }
      }
    }
    if (!any) {
      if (ld_volatile_shared(&sh->epi_done) == (uint32_t)kEpiWarps) {
        fence_cta();
        bool drained = true;
#pragma unroll
        for (int k = 0; k < kEpiWarps / kRescoreWarps; ++k)
          if (ld_volatile_shared(&sh->mail_head[r + k * kRescoreWarps]) != tail[k]) drained = false;
        if (drained) break;  // epi_done is bumped only after that warp's last post is visible
      } else {
        __nanosleep(256);
      }
    }
  }
  if (cand_counter && count) atomicAdd(cand_counter, (unsigned long long)count);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kScoreThreads, 1)
score_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ pool,
                     const WorkItem* __restrict__ items, uint32_t n_items, const PairMeta* __restrict__ pairs,
                     TopTwo* __restrict__ acc, int min_score, unsigned long long* cand_counter, uint32_t dbg) {
  // dbg (bring-up timing experiments only, results become meaningless): 1 = epilogue releases tiles unread,
  // 2 = B tiles are not loaded, 4 = hits are not posted, 8 = hits are not rescored, 16 = hits are not staged
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024 B alignment
  const uint32_t smem_a = smem0;
  const uint32_t smem_b = smem0 + kAStages * kABytes;
  const uint32_t smem_stage = smem_b + kStages * kBBytes;
  ScoreShared* sh = reinterpret_cast<ScoreShared*>(smem_raw + (smem0 - ptx::smem_u32(smem_raw)) + kAStages * kABytes +
                                                   kStages * kBBytes + kRescoreWarps * kStageSlots * kHitBytes);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();       // 0 = leader (issues the MMAs), 1 = peer
  const uint32_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap);
    for (int s = 0; s < kAStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sh->a_full[s]), 1);   // leader's producer arrives; both CTAs' bytes count
      ptx::mbar_init(ptx::smem_u32(&sh->a_empty[s]), 1);  // multicast commit
    }
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sh->b_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&sh->b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sh->t_full[s]), 1);               // multicast commit
      ptx::mbar_init(ptx::smem_u32(&sh->t_empty[s]), 2 * kEpiWarps);  // epilogue warps of BOTH CTAs (leader's copy is used)
    }
    for (int r = 0; r < kRescoreWarps; ++r)
      for (int s = 0; s < kStageSlots; ++s) ptx::mbar_init(ptx::smem_u32(&sh->h_full[r][s]), 1);
    for (int e = 0; e < kEpiWarps; ++e) {
      sh->mail_head[e] = 0;
      sh->mail_tail[e] = 0;
    }
    sh->epi_done = 0;
    ptx::fence_barrier_init();
  }
  ptx::cluster_sync();  // both CTAs resident and their barriers initialised before anything crosses the pair
  if (warp == 2) {      // one warp per CTA: paired TMEM allocation (all 512 columns: two 256-column accumulators)
    ptx::tmem_alloc_512_2sm(ptx::smem_u32(&sh->tmem_base));
    ptx::tmem_relinquish_2sm();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one lane, both CTAs)
    if (lane == 0) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      for (uint32_t it = pair_id; it < n_items; it += n_pairs) {
        const WorkItem w = items[it];
        ptx::mbar_wait(ptx::smem_u32(&sh->a_empty[as]), aph ^ 1);
        const uint32_t afull = ptx::smem_u32(&sh->a_full[as]);
        if (rank == 0) ptx::mbar_arrive_expect_tx(afull, 2 * w.m_tiles * (kABytes / 2));  // both CTAs' bytes
        for (uint32_t mh = 0; mh < w.m_tiles; ++mh)
          ptx::tma_load_2d_2sm(smem_a + as * kABytes + mh * (kABytes / 2), &tmap, afull, 0,
                               (int32_t)(w.a_row + mh * 2 * kMTile + rank * kMTile));
        if (++as == kAStages) { as = 0; aph ^= 1; }
        for (uint32_t t = 0; t < w.n_btiles; ++t) {
          ptx::mbar_wait(ptx::smem_u32(&sh->b_empty[bs]), bph ^ 1);
          const uint32_t full = ptx::smem_u32(&sh->b_full[bs]);
          if (dbg & 2) {
            if (rank == 0) ptx::mbar_arrive(full);
          } else {
            if (rank == 0) ptx::mbar_arrive_expect_tx(full, 2 * kBBytes);
            ptx::tma_load_2d_2sm(smem_b + bs * kBBytes, &tmap, full, 0,
                                 (int32_t)(w.b_row + t * kTileCols + rank * (kTileCols / 2)));
          }
          if (++bs == kStages) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one lane of the leader CTA)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_u8u8s32(2 * kMTile, kTileCols);
      uint32_t as = 0, aph = 0, bs = 0, bph = 0, ts = 0, tph = 0;
      for (uint32_t it = pair_id; it < n_items; it += n_pairs) {
        const uint32_t n_btiles = items[it].n_btiles, m_tiles = items[it].m_tiles;
        ptx::mbar_wait(ptx::smem_u32(&sh->a_full[as]), aph);
        const uint64_t adesc0 = ptx::make_kmajor_sw128_desc(smem_a + as * kABytes);
        for (uint32_t t = 0; t < n_btiles; ++t) {
          ptx::mbar_wait(ptx::smem_u32(&sh->b_full[bs]), bph);
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(smem_b + bs * kBBytes);
          for (uint32_t mh = 0; mh < m_tiles; ++mh) {
            ptx::mbar_wait(ptx::smem_u32(&sh->t_empty[ts]), tph ^ 1);
            ptx::tcgen05_fence_after();
            const uint64_t adesc = adesc0 + mh * ((kABytes / 2) >> 4);
            const uint32_t d = tmem_base + ts * kTileCols;
#pragma unroll
            for (uint32_t k = 0; k < kDim / 32; ++k)  // UMMA K = 32 bytes; advance inside the swizzle atom
              ptx::umma_i8_2sm(d, adesc + k * 2, bdesc + k * 2, idesc, k);
            ptx::umma_commit_2sm(ptx::smem_u32(&sh->t_full[ts]), 3);
            if (++ts == 2) { ts = 0; tph ^= 1; }
          }
          ptx::umma_commit_2sm(ptx::smem_u32(&sh->b_empty[bs]), 3);
          if (++bs == kStages) { bs = 0; bph ^= 1; }
        }
        ptx::umma_commit_2sm(ptx::smem_u32(&sh->a_empty[as]), 3);
        if (++as == kAStages) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ------------------------------------------------------------ filter epilogue (16 warps, both CTAs)
    const uint32_t e = warp - 4;
    const uint32_t quarter = warp & 3;            // TMEM lanes [32*quarter, +32) are visible to this warp
    const uint32_t col0 = (e >> 2) * kEpiCols;    // this warp's 64 of the tile's 256 columns
    const uint32_t lane_addr = (quarter * 32u) << 16;
    uint32_t ts = 0, tph = 0, mail_head = 0;
    for (uint32_t it = pair_id; it < n_items; it += n_pairs) {
      const WorkItem w = items[it];
      const PairMeta pm = pairs[w.pair];
      const uint32_t a_row = w.a_row + rank * kMTile + quarter * 32 + lane;  // pool row of this thread's descriptor (mh = 0)
      const uint32_t row_slot = pm.acc_off + (a_row - pm.a_row0);            // its accumulator slot
      const uint32_t col_slot0 = pm.acc_off + pm.n1 + col0;
      for (uint32_t t = 0; t < w.n_btiles; ++t) {
        for (uint32_t mh = 0; mh < w.m_tiles; ++mh) {
          ptx::mbar_wait(ptx::smem_u32(&sh->t_full[ts]), tph);
          ptx::tcgen05_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + ts * kTileCols + col0;
          int mc0 = 0, mc1 = 0;
          uint32_t v0[32], v1[32];  // both runs in flight; ptxas tracks each load's registers
          if (dbg & 1) {
#pragma unroll
            for (int x = 0; x < 32; ++x) v0[x] = v1[x] = 0;
          } else {
            ptx::tmem_ld_32x32b_x32(taddr, v0);
            ptx::tmem_ld_32x32b_x32(taddr + kRunCols, v1);
            ptx::tmem_wait_ld();
            mc0 = max_tree32(v0);
            mc1 = max_tree32(v1);
          }
          // the accumulator values are in registers: hand the TMEM buffer back to the leader's MMA warp
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t bar = ptx::smem_u32(&sh->t_empty[ts]);
            if (rank == 0) ptx::mbar_arrive(bar); else ptx::mbar_arrive_cluster(bar, 0);
          }
          if (++ts == 2) { ts = 0; tph ^= 1; }
          if (__any_sync(0xffffffffu, max(mc0, mc1) >= min_score) && !(dbg & 4)) {
            // rare: some run holds a survivor.  The exact score is the run maximum already in registers;
            // only a run with SEVERAL survivors has to be re-scored by the insert warps.
            const uint32_t arow = a_row + mh * 2 * kMTile, rslot = row_slot + mh * 2 * kMTile;
            const uint32_t brow = w.b_row + t * kTileCols + col0, cslot = col_slot0 + t * kTileCols;
            if (__any_sync(0xffffffffu, mc0 >= min_score)) {
              const uint32_t mask = survivor_mask(v0, min_score);
              const uint4 rec = __popc(mask) == 1 ? make_uint4(rslot, cslot + __ffs(mask) - 1, (uint32_t)mc0, kDirectTag)
                                                  : make_uint4(arow, brow, rslot, cslot);
              mail_post(sh, e, lane, mask != 0, rec, mail_head);
            }
            if (__any_sync(0xffffffffu, mc1 >= min_score)) {
              const uint32_t mask = survivor_mask(v1, min_score);
              const uint4 rec = __popc(mask) == 1
                                    ? make_uint4(rslot, cslot + kRunCols + __ffs(mask) - 1, (uint32_t)mc1, kDirectTag)
                                    : make_uint4(arow, brow + kRunCols, rslot, cslot + kRunCols);
              mail_post(sh, e, lane, mask != 0, rec, mail_head);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      fence_cta();
      atomicAdd(&sh->epi_done, 1u);
    }
  } else {
    // ------------------------------------------------------------ insert / re-score (warps 2, 3)
    insert_loop(sh, smem_stage, pool, acc, min_score, warp - 2, lane, cand_counter, dbg);
  }

  // the peer's shared memory and TMEM are read by MMAs the leader issued: leave together
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_512_2sm(tmem_base);
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
    for (uint32_t mh = 0; mh < 2 * w.m_tiles; ++mh) {  // 128-row sub-blocks of the strip
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
