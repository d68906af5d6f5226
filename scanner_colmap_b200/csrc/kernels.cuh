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

struct alignas(16) WorkItem {  // one strip (<= 256 rows) of image 1 against all of image 2
  uint32_t a_row;      // pool row of the strip
  uint32_t b_row;      // pool row of image 2
  uint32_t n_btiles;   // number of 256-column tiles
  uint32_t m_tiles;    // 1 or 2 accumulator row blocks (128 rows each) in this strip
  uint32_t row_slot0;  // accumulator slot of the strip's first row   (= acc_off + a_row - a_row0)
  uint32_t col_slot0;  // accumulator slot of image 2's first column  (= acc_off + n1)
  uint32_t pair;       // index into PairMeta (batch-local)
  uint32_t wait_ticket;  // 0, or the host-upload ticket whose copies must have landed before this item's rows are read
};
static_assert(sizeof(WorkItem) == 32, "WorkItem is two 16-byte words");

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
// Production score kernel: TMA -> swizzled smem -> tcgen05.mma kind::i8 -> TMEM -> filter -> insert
//
//   warp 0        TMA producer: A strip (2 x 128 rows, kept for the whole item) + ring of B tiles (256 rows)
//   warp 1        MMA issuer: per B tile two accumulator tiles (strip rows 0-127 / 128-255), 4 x K32 each,
//                 ping-pong between the two 256-column halves of TMEM
//   warps 4-11    filter epilogue: two tcgen05.ld.x64 + 3-input max tree over each thread's four 32-column runs;
//                 the accumulator buffer is released as soon as it has been read.  A run whose maximum
//                 reaches min_score is copied, exact scores and all, from registers to a shared-memory
//                 mailbox (a few vector stores; rare).
//   warps 2,3     insert: lane l examines column l of every posted run; scores >= min_score raise the best
//                 keys with fire-and-forget RED.MAX and are appended to a survivor log, from which
//                 runner_up_kernel settles the second-best keys afterwards.
//
// Where the time goes (build with -DSMB_TRACE, run tools/trace_case.py; microbenchmarks tools/smem_port.cu,
// tools/mbar_probe.cu): the kernel is bound by the epilogue warps' serial per-tile chain -- barrier wait (a try_wait
// on a complete mbarrier answers in ~58 clk, with a suspend hint in ~86), two tcgen05.ld.x64 (~170 until the buffer
// is handed back), then the max trees of the two warps that share an SMSP's ALU pipe (~230) -- not by the tensor
// pipe (538 clk per tile with these operand addresses, measured in isolation; concurrent TMA / TMEM-read / mailbox
// traffic does not slow it), and by survivor posting: without it the same kernel runs at 4.07 POP/s, with it at
// 3.31 (one warp posting delays the release of its next tile, and one of the eight posts in 85 % of the tiles).
// Variants built and measured slower: cta_group::2 CTA pairs, 16 epilogue warps, two MMA issuer threads with
// setmaxnreg warpgroups, N = 160 x 3 / N = 128 x 4 TMEM buffers, thin records + dp4a recomputation by the insert
// warps, a pre-filter split between the ALU and FMA pipes.
// =====================================================================================
constexpr int kStages = 4;                     // B-tile ring
constexpr int kAStages = 2;                    // A-strip ring (next item's strip prefetched)
constexpr int kMTile = 128;                    // UMMA M
constexpr int kABytes = 2 * kMTile * kDim;     // 32 KiB (two 128-row boxes)
constexpr int kBBytes = kTileCols * kDim;      // 32 KiB (two 128-row boxes)
#ifndef SMB_IDLE_NS
#define SMB_IDLE_NS 200
#endif
#ifndef SMB_EPI_WARPS
#define SMB_EPI_WARPS 8
#endif
constexpr int kEpiWarps = SMB_EPI_WARPS;       // 2 or 4 per TMEM lane quarter, each a column slice of the tile
constexpr int kEpiCols = kTileCols / (kEpiWarps / 4);  // accumulator columns per warp and tile
constexpr int kInsertWarps = 2;                // warps 2 and 3
constexpr int kInsertWarp0 = 2;
constexpr int kScoreWarps = 4 + kEpiWarps;     // 0 TMA, 1 MMA, 2-3 insert, 4.. epilogue
constexpr int kScoreThreads = 32 * kScoreWarps;
constexpr int kRunCols = 32;                   // accumulator columns per thread per tcgen05.ld
constexpr int kRunsPerWarp = kEpiCols / kRunCols;  // 2 or 4, all in flight at once
constexpr int kMailSlots = 128 / kEpiWarps;    // per epilogue warp: ring of survivor runs
static_assert(kRunsPerWarp == 4 || kRunsPerWarp == 2, "epilogue code is written for two or four 32-column runs per warp");

// A 32-column run of one accumulator row that holds at least one score >= min_score, copied out of the
// epilogue thread's registers: the exact scores themselves, no recomputation.
struct HitRun {
  uint32_t row_slot;   // accumulator slot of the row
  uint32_t col_slot0;  // accumulator slot of the run's first column
  uint32_t pad_[2];
  uint32_t v[kRunCols];
};
static_assert(sizeof(HitRun) == 144, "HitRun layout");

struct ScoreShared {
  uint64_t a_full[kAStages], a_empty[kAStages];
  uint64_t b_full[kStages], b_empty[kStages];
  uint64_t t_full[2], t_empty[2];
  uint32_t mail_head[kEpiWarps];  // runs posted by epilogue warp e (monotonic)
  uint32_t mail_tail[kEpiWarps];  // runs consumed by its insert warp (monotonic)
  uint32_t epi_done;              // epilogue warps that have finished
  uint32_t tmem_base;
  uint32_t pad_[2];
  HitRun mail[kEpiWarps][kMailSlots];
};
constexpr int kScoreSmemBytes = 1024 /*align slack*/ + kAStages * kABytes + kStages * kBBytes + (int)sizeof(ScoreShared);
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

#ifdef SMB_TRACE
__device__ uint32_t g_tr[3][64][4];  // [MMA, -, epilogue warp 0][tile index][event] clock stamps of CTA 0
__device__ __forceinline__ uint32_t tr_clock() {
  uint32_t c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c)::"memory");
  return c;
}
#endif
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(p)) : "memory");
  return v;
}
// Acquire load (CTA scope) of a shared-memory word: orders the loads that follow it without waiting for this
// thread's outstanding global stores/atomics, which a fence would (measured: ~2500 clk per insert pass).
__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(p)) : "memory");
  return v;
}
// Release store (CTA scope) of a shared-memory word: everything this thread -- and, through a preceding
// __syncwarp, every lane of its warp -- wrote before it is visible to a thread that then reads the word with
// ld.acquire.cta.  Publishes the mailbox head (SASS: MEMBAR.ALL.CTA + STS on the one publishing lane).
__device__ __forceinline__ void st_release_shared(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(ptx::smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Copy one 32-column run (exact scores, still in registers) into a mailbox entry: 9 vector stores.
__device__ __forceinline__ void store_run(ScoreShared* sh, uint32_t e, uint32_t pos, const uint32_t (&v)[32],
                                          uint32_t row_slot, uint32_t col_slot0) {
  const uint32_t dst = ptx::smem_u32(&sh->mail[e][pos % kMailSlots]);
  st_shared_v4(dst, row_slot, col_slot0, 0u, 0u);
#pragma unroll
  for (int q = 0; q < kRunCols / 4; ++q) st_shared_v4(dst + 16 + q * 16, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// Survivor log: with `log` != nullptr the insert warps never wait for an atomic.  They raise the best keys with
// fire-and-forget RED.MAX and append {row slot, column slot, score} to a global log; runner_up_kernel then
// offers every logged survivor that is not the final best to the runner-up keys.  The log is written in
// per-warp chunks (one returning atomic per kLogChunk survivors); unused chunk tails are zero-filled.  If the
// log overflows (adversarial inputs where almost every score survives) the host repeats the batch with
// log == nullptr, where every insertion uses the returning two-stage top2_insert2.
constexpr uint32_t kLogChunk = 256;

struct SurvivorLog {
  uint4* entries;               // {row slot, column slot, score, 0}; score 0 = unused
  unsigned long long* count;    // entries reserved so far (may exceed capacity: overflow)
  unsigned long long capacity;
};

// Insert warp r serves the mailboxes of epilogue warps r, r + kInsertWarps, ...  Lane l looks at column (run + l)
// of every posted run: a score >= min_score is a survivor of its row and its column.  The loop is kept short on
// purpose (one run per iteration, no batching): with the survivor log nothing here waits for global memory, and
// an earlier batched/compacting version spent ~1400 clk per pass on its own instruction stream and throttled
// the epilogue through mailbox back-pressure (measured with -DSMB_TRACE).
__device__ __forceinline__ void insert_loop(ScoreShared* sh, TopTwo* __restrict__ acc, SurvivorLog slog, int min_score,
                                            uint32_t r, uint32_t lane, unsigned long long* cand_counter, uint32_t dbg) {
  constexpr int kBoxes = kEpiWarps / kInsertWarps;
  uint32_t tail[kBoxes];
#pragma unroll
  for (int k = 0; k < kBoxes; ++k) tail[k] = 0;
  uint32_t count = 0;
  unsigned long long chunk_pos = 0;  // next free entry of this warp's current log chunk
  uint32_t chunk_left = 0;
  const uint32_t lt_mask = (1u << lane) - 1u;
#ifdef SMB_TRACE
  uint32_t tr_busy = 0, tr_passes = 0, tr_runs = 0, tr_idle_polls = 0;
  const uint32_t tr_start = tr_clock();
#endif
  for (;;) {
#ifdef SMB_TRACE
    const uint32_t tr0 = tr_clock();
#endif
    uint32_t got = 0;
#pragma unroll
    for (int k = 0; k < kBoxes; ++k) {
      const uint32_t e = r + k * kInsertWarps;
      const uint32_t head = ld_acquire_shared(&sh->mail_head[e]);
      if (tail[k] == head) continue;  // warp-uniform
      do {
        const uint32_t src = ptx::smem_u32(&sh->mail[e][tail[k] % kMailSlots]);
        uint32_t rs, cs, sc;
        asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rs), "=r"(cs) : "r"(src) : "memory");
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(sc) : "r"(src + 16 + lane * 4) : "memory");
        ++tail[k];
        ++got;
        const bool surv = (int)sc >= min_score;
        const uint32_t bal = __ballot_sync(0xffffffffu, surv);
        const uint32_t n = __popc(bal);
        if (n == 0 || (dbg & 8)) continue;  // warp-uniform (cannot happen for n unless the ring is misused)
        cs += lane;
        if (slog.entries) {
          if (chunk_left < n) {  // retire the chunk (zero its tail) and reserve a new one
            for (uint32_t x = lane; x < chunk_left; x += 32)
              if (chunk_pos + x < slog.capacity) slog.entries[chunk_pos + x] = make_uint4(0, 0, 0, 0);
            unsigned long long p = 0;
            if (lane == 0) p = atomicAdd(slog.count, (unsigned long long)kLogChunk);
            chunk_pos = __shfl_sync(0xffffffffu, p, 0);
            chunk_left = kLogChunk;
          }
          if (surv) {
            atomicMax(&acc[rs].k1, make_key(sc, cs));  // results unused: nothing to wait for
            atomicMax(&acc[cs].k1, make_key(sc, rs));
            const unsigned long long pos = chunk_pos + __popc(bal & lt_mask);
            if (pos < slog.capacity) slog.entries[pos] = make_uint4(rs, cs, sc, 0u);
          }
          chunk_pos += n;
          chunk_left -= n;
        } else if (surv) {
          top2_insert2(acc, rs, cs, sc);
        }
        count += surv;
      } while (tail[k] != head);
      // Every lane has consumed what it loaded from the entries (the ballot and the atomics above use the loaded
      // values, and a warp issues in order), so the loads have completed before the tail store below is issued:
      // the producer cannot overwrite a slot that is still being read.  A release here instead would wait for
      // this warp's outstanding global atomics (~2500 clk per pass, measured) -- hence the plain store.
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&sh->mail_tail[e]) = tail[k];
    }
    if (got) {
#ifdef SMB_TRACE
      tr_busy += tr_clock() - tr0;
      ++tr_passes;
      tr_runs += got;
#endif
    } else {
#ifdef SMB_TRACE
      ++tr_idle_polls;
#endif
      if (ld_volatile_shared(&sh->epi_done) == (uint32_t)kEpiWarps) {
        fence_cta();
        bool drained = true;
#pragma unroll
        for (int k = 0; k < kBoxes; ++k)
          if (ld_volatile_shared(&sh->mail_head[r + k * kInsertWarps]) != tail[k]) drained = false;
        if (drained) break;  // epi_done is bumped only after that warp's last post is visible
      } else {
        __nanosleep(SMB_IDLE_NS);
      }
    }
  }
  if (slog.entries)
    for (uint32_t x = lane; x < chunk_left; x += 32)
      if (chunk_pos + x < slog.capacity) slog.entries[chunk_pos + x] = make_uint4(0, 0, 0, 0);
  if (cand_counter && count) atomicAdd(cand_counter, (unsigned long long)count);
#ifdef SMB_TRACE
  if (blockIdx.x == 0 && lane == 0)
    printf("INS %u total %u busy %u passes %u runs %u idle polls %u | clk per pass %.1f runs per pass %.2f\n", r, tr_clock() - tr_start,
           tr_busy, tr_passes, tr_runs, tr_idle_polls, (float)tr_busy / (tr_passes ? tr_passes : 1), (float)tr_runs / (tr_passes ? tr_passes : 1));
#endif
}

// Second stage of the logged insertion: every survivor that is not the final best of its row (column)
// competes for that row's (column's) runner-up key.  Keys are distinct, so "not the best" == "key != k1".
// (128 threads, <= 32 registers: small enough to share an SM with a resident score CTA, see decide_kernel)
__global__ void __launch_bounds__(128, 16)
runner_up_kernel(const uint4* __restrict__ entries, const unsigned long long* __restrict__ count,
                 unsigned long long* __restrict__ overflow_flag, unsigned long long capacity, TopTwo* __restrict__ acc) {
  unsigned long long n = *count;
  if (n > capacity) {  // overflow: the host discards this attempt and repeats it without the log
    if (blockIdx.x == 0 && threadIdx.x == 0) *overflow_flag = 1ull;
    n = capacity;
  }
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint4 e = entries[i];
    if (e.z == 0) continue;
    const unsigned long long kr = make_key(e.z, e.y), kc = make_key(e.z, e.x);
    if (acc[e.x].k1 != kr) atomicMax(&acc[e.x].k2, kr);
    if (acc[e.y].k1 != kc) atomicMax(&acc[e.y].k2, kc);
  }
}

__global__ void __launch_bounds__(kScoreThreads, 1)
score_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, const WorkItem* __restrict__ items, uint32_t n_items,
                     const PairMeta* __restrict__ pairs, TopTwo* __restrict__ acc, SurvivorLog slog, int min_score,
                     unsigned long long* cand_counter, uint32_t dbg, unsigned long long* __restrict__ cta_busy_ns,
                     const unsigned long long* landed) {
  // landed: device word holding the newest upload ticket whose copies have completed (written by the copy engine,
  // in stream order behind the ticket's descriptor copies).  An item that names rows of a still pending HOST upload
  // carries that ticket; the TMA producer waits for it right before loading the item -- so one launch covers pairs
  // whose images are still crossing PCIe, instead of one score / runner-up / decide round per upload.  (Only copy-
  // engine work is ever waited for: it needs no SM, so the persistent CTAs cannot starve it.)
  // cta_busy_ns (profiling, may be null): [blockIdx.x] = nanoseconds this persistent CTA was busy (load balance).
  // dbg (bring-up timing experiments only, results become meaningless):
  // 2 = B tiles are not loaded, 4 = survivor runs are not posted, 8 = survivors are not inserted
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024 B alignment
  const uint32_t smem_a = smem0;
  const uint32_t smem_b = smem0 + kAStages * kABytes;
  ScoreShared* sh =
      reinterpret_cast<ScoreShared*>(smem_raw + (smem0 - ptx::smem_u32(smem_raw)) + kAStages * kABytes + kStages * kBBytes);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  unsigned long long busy_t0 = 0;
  if (cta_busy_ns && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(busy_t0));

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
    for (int e = 0; e < kEpiWarps; ++e) {
      sh->mail_head[e] = 0;
      sh->mail_tail[e] = 0;
    }
    sh->epi_done = 0;
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
      uint32_t landed_seen = 0;
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = items[it];
        if (w.wait_ticket && (int32_t)(w.wait_ticket - landed_seen) > 0) {  // rows of an upload still in flight
          unsigned long long t0 = 0, now;
          for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(landed) : "memory");
            landed_seen = (uint32_t)v;
            if ((int32_t)(w.wait_ticket - landed_seen) <= 0) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            if (now - t0 > 20000000000ull) __trap();  // 20 s: the upload never came -- fail the launch, never hang the box
            __nanosleep(500);
          }
          ptx::fence_proxy_async_global();  // the TMA (async proxy) reads what the acquire made visible
        }
        ptx::mbar_wait(ptx::smem_u32(&sh->a_empty[as]), aph ^ 1);
        const uint32_t afull = ptx::smem_u32(&sh->a_full[as]);
        ptx::mbar_arrive_expect_tx(afull, w.m_tiles * (kABytes / 2));
        for (uint32_t mh = 0; mh < w.m_tiles; ++mh)
          ptx::tma_load_2d(smem_a + as * kABytes + mh * (kABytes / 2), &tmap, afull, 0, (int32_t)(w.a_row + mh * kMTile));
        if (++as == kAStages) { as = 0; aph ^= 1; }
        for (uint32_t t = 0; t < w.n_btiles; ++t) {
          ptx::mbar_wait(ptx::smem_u32(&sh->b_empty[bs]), bph ^ 1);
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
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one lane)
    // The tensor pipe queues only an instruction or two ahead of the issuing thread (measured: ~50 extra
    // instructions between groups of four MMAs cost 28 % of the MMA rate), so the loop is software-pipelined:
    // the barrier waits for the NEXT tile (already-complete try_waits still cost ~90 clk each) are done right
    // after the current tile's MMAs have been issued, while those execute.
    if (lane == 0 && blockIdx.x < n_items) {
      constexpr uint32_t idesc = ptx::make_idesc_u8u8s32(kMTile, kTileCols);
      uint32_t as = 0, aph = 0, bs = 0, bph = 0, ts = 0, tph = 0;
      uint32_t it = blockIdx.x, t = 0, mh = 0;
      uint32_t n_btiles = items[it].n_btiles, m_tiles = items[it].m_tiles;
      uint32_t nxt_nb = 0, nxt_mt = 0;  // next item's shape, fetched a whole item ahead of its use
      if (it + gridDim.x < n_items) {
        nxt_nb = items[it + gridDim.x].n_btiles;
        nxt_mt = items[it + gridDim.x].m_tiles;
      }
      ptx::mbar_wait(ptx::smem_u32(&sh->a_full[as]), aph);
      ptx::mbar_wait(ptx::smem_u32(&sh->b_full[bs]), bph);
      ptx::mbar_wait(ptx::smem_u32(&sh->t_empty[ts]), tph ^ 1);
      ptx::tcgen05_fence_after();
      uint64_t adesc0 = ptx::make_kmajor_sw128_desc(smem_a + as * kABytes);
      uint64_t bdesc = ptx::make_kmajor_sw128_desc(smem_b + bs * kBBytes);
#ifdef SMB_TRACE
      uint32_t tr_tiles = 0;
#endif
      for (;;) {
#ifdef SMB_TRACE
        const uint32_t tr0 = tr_clock();
#endif
        const uint64_t adesc = adesc0 + mh * ((kABytes / 2) >> 4);
        const uint32_t d = tmem_base + ts * kTileCols;
#pragma unroll
        for (uint32_t k = 0; k < kDim / 32; ++k)  // UMMA K = 32 bytes; advance inside the swizzle atom
          ptx::umma_i8(d, adesc + k * 2, bdesc + k * 2, idesc, k);
        ptx::umma_commit(ptx::smem_u32(&sh->t_full[ts]));
        if (++ts == 2) { ts = 0; tph ^= 1; }
#ifdef SMB_TRACE
        if (blockIdx.x == 0 && tr_tiles < 64) {
          g_tr[0][tr_tiles][0] = tr0;         // first MMA of the tile handed to the tensor pipe
          g_tr[0][tr_tiles][1] = tr_clock();  // commit accepted
        }
        ++tr_tiles;
#endif
        // ---- everything below overlaps the execution of the MMAs just issued
        if (++mh == m_tiles) {
          mh = 0;
          ptx::umma_commit(ptx::smem_u32(&sh->b_empty[bs]));
          if (++bs == kStages) { bs = 0; bph ^= 1; }
          if (++t == n_btiles) {
            t = 0;
            ptx::umma_commit(ptx::smem_u32(&sh->a_empty[as]));
            if (++as == kAStages) { as = 0; aph ^= 1; }
            it += gridDim.x;
            if (it >= n_items) break;
            n_btiles = nxt_nb;
            m_tiles = nxt_mt;
            if (it + gridDim.x < n_items) {
              nxt_nb = items[it + gridDim.x].n_btiles;
              nxt_mt = items[it + gridDim.x].m_tiles;
            }
            ptx::mbar_wait(ptx::smem_u32(&sh->a_full[as]), aph);
            adesc0 = ptx::make_kmajor_sw128_desc(smem_a + as * kABytes);
          }
          ptx::mbar_wait(ptx::smem_u32(&sh->b_full[bs]), bph);
          bdesc = ptx::make_kmajor_sw128_desc(smem_b + bs * kBBytes);
        }
        ptx::mbar_wait(ptx::smem_u32(&sh->t_empty[ts]), tph ^ 1);
        ptx::tcgen05_fence_after();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ filter epilogue (8 warps)
    const uint32_t e = warp - 4;
    const uint32_t quarter = warp & 3;          // TMEM lanes [32*quarter, +32) are visible to this warp
    const uint32_t col0 = (e >> 2) * kEpiCols;  // which 128 of the tile's 256 columns
    const uint32_t lane_addr = (quarter * 32u) << 16;
    uint32_t ts = 0, tph = 0, mail_head = 0, mail_tail_seen = 0;
#ifdef SMB_TRACE
    uint32_t tr_wait = 0, tr_ld = 0, tr_rel = 0, tr_tree = 0, tr_post = 0, tr_tiles = 0, tr_posts = 0, tr_bp = 0, tr_bpn = 0, tr_store = 0, tr_pub = 0, tr_postonly = 0;
    const uint32_t tr_start = tr_clock();
#endif
    WorkItem w_next = blockIdx.x < n_items ? items[blockIdx.x] : WorkItem{};
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const WorkItem w = w_next;
      // the next item's descriptor is fetched a whole item ahead (everything this warp needs is in the WorkItem
      // itself, so there is no dependent PairMeta load on the tile pipeline's critical warp)
      if (it + gridDim.x < n_items) w_next = items[it + gridDim.x];
      // accumulator slot of this thread's row (mh = 0) and of this warp's first column
      const uint32_t row_slot = w.row_slot0 + quarter * 32 + lane;
      const uint32_t col_slot0 = w.col_slot0 + col0;
      for (uint32_t t = 0; t < w.n_btiles; ++t) {
        for (uint32_t mh = 0; mh < w.m_tiles; ++mh) {
#ifdef SMB_TRACE
          const uint32_t tr0 = tr_clock();
#endif
          // The accumulator is usually complete long before this warp asks.  A plain try_wait answers in ~58 clk, the
          // one with a suspend-time hint inside mbar_wait in ~86 (tools/mbar_probe.cu): ask plainly once first.
          // (+3 % on the whole kernel; the same on the MMA / TMA threads' waits, which usually do have to wait, costs 1 %.)
          if (!ptx::mbar_try_wait(ptx::smem_u32(&sh->t_full[ts]), tph)) ptx::mbar_wait(ptx::smem_u32(&sh->t_full[ts]), tph);
          ptx::tcgen05_fence_after();
#ifdef SMB_TRACE
          const uint32_t tr1 = tr_clock();
          tr_wait += tr1 - tr0;
          ++tr_tiles;
#endif
          const uint32_t taddr = tmem_base + lane_addr + ts * kTileCols + col0;
          uint32_t v0[32], v1[32], v2[32], v3[32];  // the warp's 128 columns of the tile: 128 registers per thread
          if constexpr (kRunsPerWarp == 4) {
            ptx::tmem_ld_32x32b_x64(taddr, v0, v1);
            ptx::tmem_ld_32x32b_x64(taddr + 2 * kRunCols, v2, v3);
          } else {
            ptx::tmem_ld_32x32b_x32(taddr, v0);
            ptx::tmem_ld_32x32b_x32(taddr + kRunCols, v1);
          }
          ptx::tmem_wait_ld();
#ifdef SMB_TRACE
          const uint32_t tr2 = tr_clock();
          tr_ld += tr2 - tr1;
#endif
          // every accumulator value of the tile is in registers: hand the TMEM buffer back to the MMA warp at
          // once -- the hand-off latency (commit -> wake -> read -> arrive -> wake), not the arithmetic, is what
          // the two TMEM buffers have to cover
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&sh->t_empty[ts]));
          if (++ts == 2) { ts = 0; tph ^= 1; }
#ifdef SMB_TRACE
          const uint32_t tr3 = tr_clock();
          tr_rel += tr3 - tr2;
#endif
          const int mc0 = max_tree32(v0), mc1 = max_tree32(v1);
          int mc2 = -1, mc3 = -1;  // below every score and every min_score (>= 0)
          if constexpr (kRunsPerWarp == 4) {
            mc2 = max_tree32(v2);
            mc3 = max_tree32(v3);
          }
          uint32_t lanes = __ballot_sync(0xffffffffu, __vimax3_s32(mc0, mc1, max(mc2, mc3)) >= min_score);
#ifdef SMB_TRACE
          const uint32_t tr4 = tr_clock();
          tr_tree += tr4 - tr3;
          if (lanes && !(dbg & 4)) ++tr_posts;
#endif
          if (lanes && !(dbg & 4)) {
            const uint32_t hm = (mc0 >= min_score ? 1u : 0u) | (mc1 >= min_score ? 2u : 0u) | (mc2 >= min_score ? 4u : 0u) |
                                (mc3 >= min_score ? 8u : 0u);
            const uint32_t rslot = row_slot + mh * kMTile, cslot = col_slot0 + t * kTileCols;
            do {
              const int src = __ffs(lanes) - 1;
              lanes &= lanes - 1;
              const uint32_t m = __shfl_sync(0xffffffffu, hm, src);
              const uint32_t need = __popc(m);
              if (mail_head + need - mail_tail_seen > (uint32_t)kMailSlots) {  // ring (seems) full: back-pressure
                __syncwarp();  // publish what has been written so far, or the consumer could never make room
                if (lane == 0) st_release_shared(&sh->mail_head[e], mail_head);
                uint32_t spins = 0;
#ifdef SMB_TRACE
                const uint32_t trb = tr_clock();
#endif
                do {
                  mail_tail_seen = ld_volatile_shared(&sh->mail_tail[e]);
                  if (++spins > (1u << 28)) __trap();
                } while (mail_head + need - mail_tail_seen > (uint32_t)kMailSlots);
#ifdef SMB_TRACE
                tr_bp += tr_clock() - trb;
                ++tr_bpn;
#endif
              }
#ifdef SMB_TRACE
              const uint32_t trs0 = tr_clock();
#endif
              if (lane == (uint32_t)src) {
                uint32_t pos = mail_head;
                if (m & 1u) store_run(sh, e, pos++, v0, rslot, cslot);
                if (m & 2u) store_run(sh, e, pos++, v1, rslot, cslot + kRunCols);
                if constexpr (kRunsPerWarp == 4) {
                  if (m & 4u) store_run(sh, e, pos++, v2, rslot, cslot + 2 * kRunCols);
                  if (m & 8u) store_run(sh, e, pos++, v3, rslot, cslot + 3 * kRunCols);
                }
              }
              mail_head += need;
#ifdef SMB_TRACE
              __syncwarp();
              tr_store += tr_clock() - trs0;
#endif
            } while (lanes);
#ifdef SMB_TRACE
            const uint32_t trs1 = tr_clock();
#endif
            // Release-publish: __syncwarp orders every lane's payload stores before lane 0's st.release.cta, which
            // pairs with the insert warp's ld.acquire.cta of the head (measured in bench.py: 3.15 POP/s with it, 3.14 with
            // the plain store of round 1 -- the posting warp already waits longer for its vector stores to issue).
            __syncwarp();
            if (lane == 0) st_release_shared(&sh->mail_head[e], mail_head);
#ifdef SMB_TRACE
            tr_pub += tr_clock() - trs1;
            tr_postonly += tr_clock() - tr4;
#endif
          }
#ifdef SMB_TRACE
          tr_post += tr_clock() - tr4;
          if (blockIdx.x == 0 && e == 0 && lane == 0 && tr_tiles - 1 < 64) {
            g_tr[2][tr_tiles - 1][0] = tr0;
            g_tr[2][tr_tiles - 1][1] = tr1;
            g_tr[2][tr_tiles - 1][2] = tr3;
            g_tr[2][tr_tiles - 1][3] = tr_clock();
          }
#endif
        }
      }
    }
#ifdef SMB_TRACE
    if (blockIdx.x == 0 && lane == 0)
      printf("EPI %2u tiles %u posts %u total %u | per tile: total %.1f wait %.1f ld %.1f rel %.1f tree %.1f post %.1f | ring checks %u, clk in them %u | per post: all %.1f store %.1f publish %.1f\n", e, tr_tiles,
             tr_posts, tr_clock() - tr_start, (float)(tr_clock() - tr_start) / tr_tiles, (float)tr_wait / tr_tiles, (float)tr_ld / tr_tiles,
             (float)tr_rel / tr_tiles, (float)tr_tree / tr_tiles, (float)tr_post / tr_tiles, tr_bpn, tr_bp, (float)tr_postonly / (tr_posts ? tr_posts : 1), (float)tr_store / (tr_posts ? tr_posts : 1), (float)tr_pub / (tr_posts ? tr_posts : 1));
#endif
    __syncwarp();
    if (lane == 0) {
      fence_cta();
      atomicAdd(&sh->epi_done, 1u);
    }
  } else {
    // ------------------------------------------------------------ insert (warps 2, 3)
    insert_loop(sh, acc, slog, min_score, warp - kInsertWarp0, lane, cand_counter, dbg);
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
#ifdef SMB_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence();
    const uint32_t t0 = g_tr[2][32][0];
    for (int q = 32; q < 44; ++q)
      printf("tile %2d buf %d | MMA issue %6d..%6d | EPI wait-from %6d t_full %6d released %6d done %6d\n", q, q & 1,
             (int)(g_tr[0][q][0] - t0), (int)(g_tr[0][q][1] - t0), (int)(g_tr[2][q][0] - t0), (int)(g_tr[2][q][1] - t0),
             (int)(g_tr[2][q][2] - t0), (int)(g_tr[2][q][3] - t0));
  }
#endif
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_512(tmem_base);
  }
  if (cta_busy_ns && threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    cta_busy_ns[blockIdx.x] = t1 - busy_t0;
  }
}

#ifdef SMB_TEST_ENGINES
// Test-only device cross-check (compiled only with -DSMB_TEST_ENGINES, i.e. into the tests' libsmb_test.so, never
// into the product libsmb.so): the same contract on CUDA cores (__dp4a), no tensor cores, no TMA.
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

#endif  // SMB_TEST_ENGINES

// Plan upload without the copy engine: the per-call plan (pair metas, work items) sits in pinned host memory,
// which the device reads directly (UVA).  A cudaMemcpyAsync would queue behind whatever descriptor uploads
// are already in the host->device copy engine's FIFO and stall the score kernel it feeds.
__global__ void __launch_bounds__(256)
fetch_words_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src_host, size_t n_words) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src_host[i];
}

// =====================================================================================
// decide: FindBestMatchesOneWay tests + cross-check + ordered compaction, one CTA per pair
// =====================================================================================
// Two CTA shapes.  128 threads and <= 32 registers per thread: a score CTA leaves 4,096 registers, ~10 KB of shared
// memory and 1,664 thread slots of its SM unused, exactly enough for ONE such CTA -- so the runner-up and decide
// kernels of one sub-batch (second stream) run underneath the score kernel of the next one.  512 threads: for the
// sub-batch whose decide runs alone (the last one), where more threads per pair hide more latency.
constexpr int kDecideThreadsSmall = 128;
constexpr int kDecideThreadsLarge = 512;

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

// `out` and `pair_out` point into PINNED HOST memory (zero-copy): matches cross PCIe as the kernel produces them,
// coalesced per warp (lanes holding a match write consecutive 8-byte entries), so there is no device staging
// buffer, no device-to-host copy after the kernel and no host round trip to learn how much to copy.  If the
// matches of this pair would not fit into `out_cap` entries the pair is skipped and *overflow set: the host then
// repeats the call with a worst-case sized buffer.
template <int kDecideThreads>
__global__ void __launch_bounds__(kDecideThreads, kDecideThreads == kDecideThreadsSmall ? 16 : 1)
decide_kernel(const PairMeta* __restrict__ pairs, TopTwo* __restrict__ acc, const float* __restrict__ lut,
              float max_ratio, float max_distance, int cross_check, uint2* __restrict__ out /* FeatureMatch */,
              unsigned long long out_cap, unsigned long long* __restrict__ out_total,
              unsigned long long* __restrict__ overflow, PairOut* __restrict__ pair_out) {
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
    const bool fits = base + total <= out_cap;
    if (!fits) *overflow = 1ull;
    s_base = fits ? static_cast<uint32_t>(base) : 0xFFFFFFFFu;
    pair_out[pm.out_slot] = PairOut{static_cast<uint32_t>(base), fits ? total : 0u};
  }
  __syncthreads();
  uint32_t running = s_base;
  // Whatever happens below, this CTA leaves its pair's accumulator slots zeroed (after its last read of them): the
  // next sub-batch / call that reuses the region then needs no 16-byte-per-slot memset of its own -- the stores hide
  // under this kernel's PCIe-bound match writes.
  auto clear_slots = [&]() {
    __syncthreads();
    for (uint32_t x = tid; x < pm.n1 + pm.n2; x += kDecideThreads) rows[x] = TopTwo{0ull, 0ull};
  };
  if (running == 0xFFFFFFFFu) {  // whole CTA: this attempt is being discarded
    clear_slots();
    return;
  }

  // pass 2: ordered write (ascending idx1), block-wide exclusive scan per chunk of rows.  The matches of a chunk are
  // compacted into shared memory first and flushed in dense runs: every lane of a warp then stores 8 consecutive
  // bytes (256 contiguous bytes per warp store), which crosses PCIe as full lines -- writing each match from the
  // thread that decided it (a quarter of the lanes active, 32-byte fragments) reached only ~31 GB/s.
  __shared__ uint2 sbuf[2 * kDecideThreads];
  uint32_t buffered = 0;
  for (uint32_t i0 = 0; i0 < pm.n1; i0 += kDecideThreads) {
    const uint32_t i = i0 + tid;
    int m12 = -1;
    if (i < pm.n1) m12 = static_cast<int>(static_cast<uint32_t>(rows[i].k1));
    const uint32_t ballot = __ballot_sync(0xffffffffu, m12 >= 0);
    __syncthreads();  // warp_sums reuse; the previous flush has finished reading sbuf
    if (lane == 0) warp_sums[wid] = __popc(ballot);
    __syncthreads();
    uint32_t before = 0, chunk_total = 0;
#pragma unroll
    for (int w = 0; w < kDecideThreads / 32; ++w) {
      const uint32_t s = warp_sums[w];
      if (w < (int)wid) before += s;
      chunk_total += s;
    }
    if (m12 >= 0) sbuf[buffered + before + __popc(ballot & ((1u << lane) - 1))] = make_uint2(i, static_cast<uint32_t>(m12));
    buffered += chunk_total;
    if (buffered >= (uint32_t)kDecideThreads || i0 + kDecideThreads >= pm.n1) {  // uniform across the CTA
      __syncthreads();
      for (uint32_t x = tid; x < buffered; x += kDecideThreads) out[running + x] = sbuf[x];
      running += buffered;
      buffered = 0;
    }
  }
  clear_slots();
}

}  // namespace smb
