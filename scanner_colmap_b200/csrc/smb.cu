// Host side of libsmb: the C ABI of include/smb.h.
//
// Replaces, for the Scanner op SequentialMatchingCPU, the call
//   colmap::MatchSiftFeaturesCPU(...)      /root/reference/integration/op_cpp/sequential_matching.cc:154
// and the per-row descriptor re-deserialisation feeding it
//   read_matrix_from_element<FeatureDescriptors>   io.cc:181-194, sequential_matching.cc:120-121
// with a per-handle descriptor pool in HBM and batched pair matching on one B200.
//
// There is deliberately no CPU path in this file: every entry point that computes needs an
// sm_100 device and fails with SMB_ENODEVICE / SMB_ECUDA otherwise.
#include "../../include/smb.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "verify.cuh"

namespace {

using namespace smb;

thread_local std::string g_create_error;

struct ImageEntry {
  uint32_t row0;   // first pool row
  uint32_t n;      // descriptor count
  uint32_t rows;   // reserved pool rows (multiple of kRowPad)
  uint64_t up_seq; // 0 = resident; otherwise the upload ticket (smb_put_images_async) that fills it
  bool has_kp = false;  // keypoint positions present (smb_put_keypoints)
};

struct Filter {
  int32_t min_best;   // smallest score whose distance passes max_distance (INT32_MAX if none)
  int32_t min_score;  // smallest score that can change any decision, clamped to >= 1
};

struct Sub {  // one internal batch of a match call: pairs [first, last), its work items and accumulator slots
  size_t first, last, item0, items, acc;
  uint64_t split_ticket, wait_ticket;
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    size_t want = std::max(n, cap + cap / 2);
    T* q = nullptr;
    cudaError_t e = cudaMalloc(&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p) cudaFree(p);
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    size_t want = std::max(n, cap + cap / 2);
    T* q = nullptr;
    cudaError_t e = cudaMallocHost(&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p) cudaFreeHost(p);
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

// A result lives in pinned, device-mapped host memory that decide_kernel writes directly (zero-copy): the match
// lists, the per-pair {start, count} table (caller order) and a copy of the device counters.
struct smb_result {
  size_t npairs = 0;
  PairOut* pair_out = nullptr;  // pinned [pair_cap]
  size_t pair_cap = 0;
  smb_match* matches = nullptr;  // pinned [matches_cap]
  size_t matches_cap = 0;
  size_t matches_limit = 0;      // entries the device may write in the current attempt (<= matches_cap)
  unsigned long long* counters = nullptr;  // pinned [kNumCounters]: the device counters as of the end of the call
  size_t total = 0;
  smb_tvg* tvg = nullptr;        // pinned [tvg_cap], written by verify_kernel (smb_result_verify)
  size_t tvg_cap = 0;
  smb_match* inliers = nullptr;  // pinned [inliers_cap], same offsets as matches
  size_t inliers_cap = 0;
  // EstimateMultiple rounds (SMB_TVG_MULTIPLE_MODELS): the matches still unexplained, and one round's verdicts
  smb_match* rem_matches = nullptr;
  size_t rem_matches_cap = 0;
  PairOut* rem_po = nullptr;
  size_t rem_po_cap = 0;
  smb_tvg* round_tvg = nullptr;
  size_t round_tvg_cap = 0;
  smb_match* round_inl = nullptr;
  size_t round_inl_cap = 0;
  bool verified = false;
  // between smb_match_pairs_begin and smb_result_wait
  bool pending = false;
  bool used_log = false;
  size_t worst_case = 0;        // upper bound of the matches this call can produce
  std::vector<uint64_t> keys;   // the call's pairs, kept for the (rare) repeat after an overflow
  std::vector<cudaEvent_t> ev;  // profiling: 4 per sub-batch + 2
  size_t n_subs = 0;
  std::vector<uint8_t> sub_has_items;
  std::vector<uint32_t> sub_grid;  // CTAs of each sub-batch's score launch (0 = none)
  uint64_t ops = 0;
  uint32_t launches = 0, score_launches = 0, plan_uploaded = 0;
};

struct smb_handle {
  int device = -1;
  int num_sms = 0;
  smb_options opts{};
  float max_ratio_f = 0.f, max_distance_f = 0.f;
  Filter filter{};
  std::string err;
  cudaStream_t stream = nullptr;      // synchronous uploads, accumulator clears, score kernels
  cudaStream_t stream_b = nullptr;    // runner-up and decide kernels of sub-batch k, underneath the score kernel of k+1
  std::vector<cudaEvent_t> chain_ev;  // 2 per sub-batch: scored (main stream), decided (second stream)
  // A call of >= 256 pairs is cut into `waves` sub-batches so that the delivery of all but the last one's matches hides
  // under scoring.  Adaptive (SMB_WAVES unset): one sub-batch until a call's exposed tail -- runner-up + decide after
  // its last score kernel -- exceeded 12 % of the call (several GPUs sharing the host's I/O fabric), then four.  More
  // sub-batches mean smaller score launches (each pays its own tail: -3 % score-kernel rate at four), so a lone
  // GPU, whose tail is ~8 %, stays at one.
  int waves = 1;
  bool waves_adaptive = true;
  cudaEvent_t tail_ev[3] = {};        // call start, last score kernel done (main stream), call end (second stream)
  cudaStream_t stream_up = nullptr;   // asynchronous uploads (smb_put_images_async): copy engine under the score kernels
  static constexpr int kUpRing = 64;
  cudaEvent_t up_ev[kUpRing] = {};    // up_ev[t % kUpRing] fires when upload ticket t has landed
  uint64_t up_issued = 0;             // last ticket handed out
  uint64_t up_open = 0;               // ticket whose copies are being queued right now (0 = none)
  std::vector<uint32_t> plan_order;   // scratch of match_keys_impl
  std::vector<uint64_t> plan_ticket;
  std::vector<const ImageEntry*> plan_entries;
  uint64_t up_synced = 0;             // tickets <= this are known to have landed
  unsigned long long* d_landed = nullptr;       // device word: newest ticket whose copies have completed (see score kernel)
  unsigned long long* h_ticket_vals = nullptr;  // pinned [kUpRing]: the values copied into it
  uint8_t up_fast[kUpRing] = {};      // ticket kind: 1 = device-to-device adoption (NVLink halo: lands within
                                      // microseconds, never worth a sub-batch of its own), 0 = host upload over PCIe
  cudaEvent_t ev_ext = nullptr;  // marks the producer stream's position in smb_put_images_device_async
  cudaEvent_t ev_ext2 = nullptr; // smb_wait_stream

  // keypoint positions, one float2 per descriptor-pool row (only used by smb_result_verify)
  float2* kp_pool = nullptr;
  DevBuf<float4> d_pts;               // verification scratch: matched point pairs
  smb_result* last_result = nullptr;  // the result the device-side plan (d_pairs) currently describes
  // descriptor pool
  uint8_t* pool = nullptr;
  uint32_t pool_rows = 0;
  std::map<uint32_t, uint32_t> free_list;  // row0 -> rows (coalesced)
  std::unordered_map<uint64_t, ImageEntry> images;
  CUtensorMap tmap{};
  bool tmap_valid = false;

  // acos table
  std::vector<float> lut_host;
  float* lut_dev = nullptr;

  // per-call scratch
  DevBuf<PairMeta> d_pairs;
  DevBuf<WorkItem> d_items;
  DevBuf<TopTwo> d_acc;
  size_t acc_zero_slots = 0;  // leading slots of d_acc known to be zero: decide_kernel clears what a sub-batch dirtied
  static constexpr int kNumCounters = 8;
  unsigned long long* d_counters = nullptr;  // [0] out_total, [1] candidates, [2] survivor-log entries (log half 0),
                                             // [3] log overflowed, [4] result buffer overflowed, [5] log entries (half 1)
  DevBuf<uint4> d_log;                       // survivor log (kernels.cuh SurvivorLog)
  DevBuf<unsigned long long> d_cta_busy;     // profiling: per score launch, per CTA busy nanoseconds
  PinnedBuf<unsigned long long> h_cta_busy;
  size_t log_cap = (size_t)16 << 20;         // entries; SMB_LOG_CAP overrides (tests force the overflow path)
  PinnedBuf<PairMeta> h_pairs;   // the plan the device copy (d_pairs / d_items) was fetched from
  PinnedBuf<WorkItem> h_items;
  std::vector<PairMeta> plan_pairs;  // the plan being built; uploaded only if it differs from h_pairs / h_items
  std::vector<WorkItem> plan_items;
  size_t dev_plan_pairs = 0, dev_plan_items = 0;  // extent of the valid device copy (0 = none)
  // Steady state (the same pair list over an unchanged pool layout, e.g. every step of a resident window, or a
  // window whose halo images are refreshed in place): the whole plan of the previous call is reused, so a call costs
  // the host a 14 KB compare and five launches instead of ~0.25 ms of planning with the GPU idle.
  uint64_t layout_epoch = 1;          // bumped whenever an image appears, disappears or moves
  uint64_t plan_epoch = 0;            // layout the stored plan was made for (0 = no stored plan)
  std::vector<uint64_t> plan_keys;
  std::vector<Sub> plan_subs;
  size_t plan_out_cap = 0, plan_max_acc = 0, plan_acc_budget = 0;
  uint64_t plan_ops = 0;

  std::vector<smb_result*> result_pool;
  smb_result* inflight = nullptr;        // begun, not yet waited for
  std::vector<std::pair<uint32_t, uint32_t>> pending_free;  // rows evicted while a call was in flight
  smb_timing timing{};
  size_t acc_budget = (size_t)64 << 20;  // accumulator slots per internal batch (16 B each; must stay < 2^32)
  double matches_per_pair = 0.0;         // most matches per pair any call produced so far (sizes the result buffer)
  size_t result_cap_override = 0;        // SMB_RESULT_CAP: first-attempt capacity in matches (tests force the overflow repeat)

  PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
  uint32_t dbg_flags = 0;  // SMB_DEBUG_FLAGS: bring-up timing experiments (see score_tcgen05_kernel)
};

namespace {

int fail(smb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define SMB_CUDA(h, expr)                                                                              \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return fail((h), e__ == cudaErrorMemoryAllocation ? SMB_ENOMEM : SMB_ECUDA, "%s: %s (%s:%d)", #expr, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                        \
  } while (0)

// ------------------------------------------------------------------------------------------
// Filter derivation.  lut[d] = acosf(min(d * 2^-18, 1)) computed by THIS process' libm, i.e. the
// very function COLMAP's FindBestMatchesOneWay evaluates; the device only ever looks values up.
//
//   min_best  = min{ d : lut[d] <= max_distance }                      (best below it fails test 1)
//   M         = max{ lut[d] : lut[d] <= max_distance }                 (largest passing distance)
//   min_second= min{ s : max_ratio * lut[s] <= M }                     (second below it can never
//                                                                       make test 2 reject)
//   min_score = max(1, min(min_best, min_second))
//
// No monotonicity of lut is assumed.  Scores below min_score are dropped before the top-2
// accumulators; DESIGN.md ("Filter") proves every FeatureMatch is unchanged.
// ------------------------------------------------------------------------------------------
Filter derive_filter(const std::vector<float>& lut, float max_ratio, float max_distance) {
  Filter f;
  int64_t min_best = -1;
  float M = -1.f;
  for (int d = 0; d < kLutSize; ++d) {
    if (lut[d] <= max_distance) {
      if (min_best < 0) min_best = d;
      M = std::max(M, lut[d]);
    }
  }
  if (min_best < 0) {  // nothing can ever pass max_distance (also covers NaN thresholds)
    f.min_best = INT32_MAX;
    f.min_score = INT32_MAX;
    return f;
  }
  int64_t min_second = kLutSize;  // scores >= 512^2 share lut[512^2]; handled by the loop's last entry
  for (int s = 0; s < kLutSize; ++s) {
    const float rhs = max_ratio * lut[s];
    if (!(M < rhs)) {  // "best_normed >= max_ratio * second_normed" possible (NaN-safe)
      min_second = s;
      break;
    }
  }
  f.min_best = (int32_t)min_best;
  f.min_score = (int32_t)std::max<int64_t>(1, std::min(min_best, min_second));
  return f;
}

int apply_options(smb_handle* h, const smb_options* opts) {
  if (!opts) return fail(h, SMB_EINVAL, "options pointer is null");
#ifdef SMB_TEST_ENGINES
  if (opts->engine != SMB_ENGINE_TCGEN05 && opts->engine != SMB_ENGINE_DP4A)
    return fail(h, SMB_EINVAL, "unknown engine %d", opts->engine);
#else
  if (opts->engine != SMB_ENGINE_TCGEN05)
    return fail(h, SMB_EINVAL, "engine %d is not part of this build (the product library has the tcgen05 engine only)",
                opts->engine);
#endif
  h->opts = *opts;
  h->max_ratio_f = (float)opts->max_ratio;        // double -> float exactly where COLMAP narrows
  h->max_distance_f = (float)opts->max_distance;
  h->filter = derive_filter(h->lut_host, h->max_ratio_f, h->max_distance_f);
  return SMB_OK;
}

// ------------------------------------------------------------------------------------------
// Descriptor pool: one device allocation of pool_rows x 128 B; images occupy row ranges padded to
// kRowPad with zero rows (a zero descriptor scores 0 against everything and can never be a
// candidate), so TMA boxes never need masking.
// ------------------------------------------------------------------------------------------
int drain_uploads(smb_handle* h);

int encode_tmap(smb_handle* h) {
  h->tmap_valid = false;
  if (!h->pool_rows) return SMB_OK;
  cuuint64_t gdim[2] = {(cuuint64_t)kDim, (cuuint64_t)h->pool_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kDim};
  cuuint32_t box[2] = {(cuuint32_t)kDim, (cuuint32_t)kMTile};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = h->encode_tiled(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, h->pool, gdim, gstride, box, estride,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, SMB_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  h->tmap_valid = true;
  return SMB_OK;
}

void free_rows(smb_handle* h, uint32_t row0, uint32_t rows) {
  if (!rows) return;
  auto it = h->free_list.emplace(row0, rows).first;
  auto nx = std::next(it);
  if (nx != h->free_list.end() && it->first + it->second == nx->first) {
    it->second += nx->second;
    h->free_list.erase(nx);
  }
  if (it != h->free_list.begin()) {
    auto pv = std::prev(it);
    if (pv->first + pv->second == it->first) {
      pv->second += it->second;
      h->free_list.erase(it);
    }
  }
}

int grow_pool(smb_handle* h, uint32_t min_extra_rows) {
  uint64_t want = std::max<uint64_t>((uint64_t)h->pool_rows * 2, (uint64_t)h->pool_rows + min_extra_rows);
  want = std::max<uint64_t>(want, (uint64_t)(64u << 20) / kDim);  // start at 64 MiB
  want = (want + kRowPad - 1) / kRowPad * kRowPad;
  if (want > 0xFFFFFF00ull) return fail(h, SMB_ENOMEM, "descriptor pool would exceed 2^32 rows");
  if (int rc = drain_uploads(h)) return rc;  // pending uploads target the old allocation
  if (h->inflight) return fail(h, SMB_EINVAL, "the descriptor pool must grow while a match call is in flight: wait for it first");
  uint8_t* np = nullptr;
  SMB_CUDA(h, cudaMalloc(&np, want * kDim));
  float2* nk = nullptr;
  SMB_CUDA(h, cudaMalloc(&nk, want * sizeof(float2)));
  if (h->pool) {
    SMB_CUDA(h, cudaMemcpyAsync(np, h->pool, (size_t)h->pool_rows * kDim, cudaMemcpyDeviceToDevice, h->stream));
    SMB_CUDA(h, cudaMemcpyAsync(nk, h->kp_pool, (size_t)h->pool_rows * sizeof(float2), cudaMemcpyDeviceToDevice, h->stream));
    SMB_CUDA(h, cudaStreamSynchronize(h->stream));
    SMB_CUDA(h, cudaFree(h->pool));
    SMB_CUDA(h, cudaFree(h->kp_pool));
  }
  h->kp_pool = nk;
  h->dev_plan_pairs = h->dev_plan_items = 0;  // (the plan holds pool rows, which are unchanged, but be safe)
  ++h->layout_epoch;
  const uint32_t old_rows = h->pool_rows;
  h->pool = np;
  h->pool_rows = (uint32_t)want;
  free_rows(h, old_rows, (uint32_t)want - old_rows);
  return encode_tmap(h);
}

int alloc_rows(smb_handle* h, uint32_t rows, uint32_t* row0) {
  if (rows == 0) {
    *row0 = 0;
    return SMB_OK;
  }
  for (int attempt = 0; attempt < 2; ++attempt) {
    for (auto it = h->free_list.begin(); it != h->free_list.end(); ++it) {
      if (it->second >= rows) {
        *row0 = it->first;
        const uint32_t rest = it->second - rows, rest0 = it->first + rows;
        h->free_list.erase(it);
        if (rest) h->free_list.emplace(rest0, rest);
        return SMB_OK;
      }
    }
    int rc = grow_pool(h, rows);
    if (rc != SMB_OK) return rc;
  }
  return fail(h, SMB_ENOMEM, "descriptor pool allocation of %u rows failed", rows);
}

// Everything queued on the upload stream has landed (needed before pool rows are recycled or the pool moves).
// A ticket that is still being filled (up_open: more copies of it are about to be queued) is NOT marked as
// landed by this.
int drain_uploads(smb_handle* h) {
  const uint64_t done = h->up_open ? h->up_open - 1 : h->up_issued;
  if (h->up_synced != h->up_issued) {
    SMB_CUDA(h, cudaStreamSynchronize(h->stream_up));
    h->up_synced = std::max(h->up_synced, done);
  }
  return SMB_OK;
}

void poll_uploads(smb_handle* h) {  // tickets whose event has fired need no waiting any more
  while (h->up_synced < h->up_issued && (h->up_open == 0 || h->up_synced + 1 < h->up_open) &&
         cudaEventQuery(h->up_ev[(h->up_synced + 1) % smb_handle::kUpRing]) == cudaSuccess)
    ++h->up_synced;
  cudaGetLastError();  // cudaErrorNotReady from the query is not an error
}

// Give an image's pool rows back.  No device synchronisation: every kernel that could read them belongs to a match
// call, and a match call has either completed (smb_match_pairs / smb_result_wait return after both streams have been
// synchronised) or is the one call in flight, in which case the rows are parked until its wait.  Copies into
// recycled rows are ordered behind earlier work of the stream they are queued on; only an upload that is still
// in flight INTO these rows on the upload stream has to be waited for.
int retire_rows(smb_handle* h, const ImageEntry& e) {
  if (e.up_seq > h->up_synced) poll_uploads(h);
  if (e.up_seq > h->up_synced)
    if (int rc = drain_uploads(h)) return rc;
  if (h->inflight)
    h->pending_free.emplace_back(e.row0, e.rows);
  else
    free_rows(h, e.row0, e.rows);
  return SMB_OK;
}

int put_image_impl(smb_handle* h, uint64_t key, const void* src, size_t n, size_t d, cudaMemcpyKind kind,
                   cudaStream_t stream = nullptr, uint64_t up_seq = 0) {
  if (!stream) stream = h->stream;
  if (d != (size_t)kDim) return fail(h, SMB_EINVAL, "descriptor dimension must be 128, got %zu", d);
  if (n && !src) return fail(h, SMB_EINVAL, "descriptor pointer is null");
  if (n > 0x7FFFFFFFu) return fail(h, SMB_EINVAL, "too many descriptors in one image: %zu", n);
  auto old = h->images.find(key);
  if (old != h->images.end() && old->second.n == n && !h->inflight) {
    // Same id, same size, nothing reading the pool: refresh the rows in place.  The layout (and with it any plan
    // built on it) stays valid; only the upload ticket changes.  Copies are ordered per stream, so only a pending
    // upload on the OTHER stream into these rows has to land first.
    ImageEntry& e = old->second;
    if (e.up_seq > h->up_synced && stream != h->stream_up) {
      poll_uploads(h);
      if (e.up_seq > h->up_synced)
        if (int rc = drain_uploads(h)) return rc;
    }
    if (n) SMB_CUDA(h, cudaMemcpyAsync(h->pool + (size_t)e.row0 * kDim, src, n * kDim, kind, stream));
    e.up_seq = up_seq;
    e.has_kp = false;  // new descriptors: the old keypoint positions no longer belong to them
    return SMB_OK;
  }
  if (old != h->images.end()) {
    if (int rc = retire_rows(h, old->second)) return rc;
    h->images.erase(old);
  }
  ++h->layout_epoch;
  ImageEntry e;
  e.n = (uint32_t)n;
  e.rows = (uint32_t)((n + kRowPad - 1) / kRowPad * kRowPad);
  e.up_seq = up_seq;
  int rc = alloc_rows(h, e.rows, &e.row0);
  if (rc != SMB_OK) return rc;
  if (n) {
    uint8_t* dst = h->pool + (size_t)e.row0 * kDim;
    SMB_CUDA(h, cudaMemcpyAsync(dst, src, n * kDim, kind, stream));
    if (e.rows > n) SMB_CUDA(h, cudaMemsetAsync(dst + n * kDim, 0, (size_t)(e.rows - n) * kDim, stream));
  }
  h->images.emplace(key, e);
  return SMB_OK;
}

// Marks an upload ticket as "being filled" for the duration of a put call (see drain_uploads).  If the call
// fails half way the ticket's event is recorded anyway, so nothing can wait on a stale event.
struct OpenTicket {
  smb_handle* h;
  uint64_t ticket;
  bool recorded = false;
  OpenTicket(smb_handle* hh, uint64_t t) : h(hh), ticket(t) { h->up_open = t; }
  cudaError_t finish() {
    recorded = true;
    h->up_open = 0;
    // in stream order behind the ticket's copies: the device-visible "landed" word, then the event
    unsigned long long* v = h->h_ticket_vals + ticket % smb_handle::kUpRing;
    *v = ticket;
    cudaError_t e = cudaMemcpyAsync(h->d_landed, v, sizeof *v, cudaMemcpyHostToDevice, h->stream_up);
    cudaError_t e2 = cudaEventRecord(h->up_ev[ticket % smb_handle::kUpRing], h->stream_up);
    return e != cudaSuccess ? e : e2;
  }
  ~OpenTicket() {
    if (!recorded) finish();
  }
};

smb_result* acquire_result(smb_handle* h) {
  if (!h->result_pool.empty()) {
    smb_result* r = h->result_pool.back();
    h->result_pool.pop_back();
    return r;
  }
  smb_result* r = new (std::nothrow) smb_result();
  if (r && cudaMallocHost(&r->counters, smb_handle::kNumCounters * sizeof(unsigned long long)) != cudaSuccess) {
    cudaGetLastError();
    delete r;
    return nullptr;
  }
  return r;
}

void destroy_result(smb_result* r) {
  if (!r) return;
  if (r->matches) cudaFreeHost(r->matches);
  if (r->pair_out) cudaFreeHost(r->pair_out);
  if (r->counters) cudaFreeHost(r->counters);
  if (r->tvg) cudaFreeHost(r->tvg);
  if (r->inliers) cudaFreeHost(r->inliers);
  if (r->rem_matches) cudaFreeHost(r->rem_matches);
  if (r->rem_po) cudaFreeHost(r->rem_po);
  if (r->round_tvg) cudaFreeHost(r->round_tvg);
  if (r->round_inl) cudaFreeHost(r->round_inl);
  for (cudaEvent_t e : r->ev) cudaEventDestroy(e);
  delete r;
}

template <typename T>
bool reserve_pinned(T** p, size_t* cap, size_t need) {
  if (need <= *cap) return true;
  const size_t want = std::max<size_t>(std::max(need, *cap + *cap / 2), 1024);
  T* q = nullptr;
  if (cudaMallocHost(&q, want * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (*p) cudaFreeHost(*p);
  *p = q;
  *cap = want;
  return true;
}

}  // namespace

// ==========================================================================================
extern "C" {

void smb_default_options(smb_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);
  o->max_ratio = 0.8;       // colmap.proto:14
  o->max_distance = 0.7;    // colmap.proto:17
  o->cross_check = 1;       // colmap.proto:20
  o->max_num_matches = 32768;  // colmap.proto:23
  o->engine = SMB_ENGINE_TCGEN05;
  o->profile = 0;
}

int smb_abi_version(void) { return SMB_ABI_VERSION; }

const char* smb_last_error(const smb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int smb_create(int cuda_device, const smb_options* opts, smb_handle** out) {
  if (!out) return fail(nullptr, SMB_EINVAL, "out pointer is null");
  *out = nullptr;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(nullptr, SMB_ENODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
  if (cuda_device < 0 || cuda_device >= count)
    return fail(nullptr, SMB_EINVAL, "cuda_device %d out of range [0, %d)", cuda_device, count);
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, cuda_device);
  if (ce != cudaSuccess) return fail(nullptr, SMB_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ce));
  if (prop.major != 10)
    return fail(nullptr, SMB_ENODEVICE, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", cuda_device,
                prop.name, prop.major, prop.minor);
  smb_handle* h = new (std::nothrow) smb_handle();
  if (!h) return fail(nullptr, SMB_ENOMEM, "out of host memory");
  h->device = cuda_device;
  h->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("SMB_DEBUG_FLAGS")) h->dbg_flags = (uint32_t)strtoul(e, nullptr, 0);
  if (const char* e = getenv("SMB_LOG_CAP")) h->log_cap = (size_t)strtoull(e, nullptr, 0);
  if (const char* e = getenv("SMB_WAVES")) {
    h->waves = std::max(1, atoi(e));
    h->waves_adaptive = false;
  }
  if (const char* e = getenv("SMB_RESULT_CAP")) h->result_cap_override = (size_t)strtoull(e, nullptr, 0);
  if (const char* e = getenv("SMB_ACC_BUDGET")) h->acc_budget = std::max<size_t>(1, (size_t)strtoull(e, nullptr, 0));  // tests: force sub-batches
  int rc = SMB_OK;
  auto bail = [&](int code) {
    g_create_error = h->err;
    smb_destroy(h);
    return code;
  };
#define SMB_CUDA_C(expr)                                                                     \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      fail(h, SMB_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return bail(SMB_ECUDA);                                                                \
    }                                                                                        \
  } while (0)
  SMB_CUDA_C(cudaSetDevice(cuda_device));
  SMB_CUDA_C(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  SMB_CUDA_C(cudaStreamCreateWithFlags(&h->stream_b, cudaStreamNonBlocking));
  SMB_CUDA_C(cudaStreamCreateWithFlags(&h->stream_up, cudaStreamNonBlocking));
  for (auto& e : h->up_ev) SMB_CUDA_C(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  SMB_CUDA_C(cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming));
  SMB_CUDA_C(cudaEventCreateWithFlags(&h->ev_ext2, cudaEventDisableTiming));
  for (auto& e : h->tail_ev) SMB_CUDA_C(cudaEventCreate(&e));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SMB_CUDA_C(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      fail(h, SMB_ECUDA, "driver does not export cuTensorMapEncodeTiled");
      return bail(SMB_ECUDA);
    }
    h->encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  // the acos table: host libm, the same function the reference evaluates
  h->lut_host.resize(kLutSize);
  {
    const float kDistNorm = 1.0f / (512.0f * 512.0f);
    for (int d = 0; d < kLutSize; ++d) h->lut_host[d] = acosf(std::min(kDistNorm * (float)d, 1.0f));
  }
  SMB_CUDA_C(cudaMalloc(&h->lut_dev, kLutSize * sizeof(float)));
  SMB_CUDA_C(cudaMemcpyAsync(h->lut_dev, h->lut_host.data(), kLutSize * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  SMB_CUDA_C(cudaMalloc(&h->d_counters, smb_handle::kNumCounters * sizeof(unsigned long long)));
  SMB_CUDA_C(cudaMalloc(&h->d_landed, sizeof(unsigned long long)));
  SMB_CUDA_C(cudaMemsetAsync(h->d_landed, 0, sizeof(unsigned long long), h->stream));
  SMB_CUDA_C(cudaMallocHost(&h->h_ticket_vals, smb_handle::kUpRing * sizeof(unsigned long long)));
  SMB_CUDA_C(cudaFuncSetAttribute(score_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScoreSmemBytes));
  SMB_CUDA_C(cudaStreamSynchronize(h->stream));
#undef SMB_CUDA_C
  smb_options def;
  smb_default_options(&def);
  rc = apply_options(h, opts ? opts : &def);
  if (rc != SMB_OK) return bail(rc);
  *out = h;
  return SMB_OK;
}

void smb_destroy(smb_handle* h) {
  if (!h) return;
  if (h->device >= 0) cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream_b) cudaStreamSynchronize(h->stream_b);
  if (h->stream_up) cudaStreamSynchronize(h->stream_up);
  for (cudaEvent_t e : h->chain_ev) cudaEventDestroy(e);
  if (h->inflight) destroy_result(h->inflight);  // begun and never waited for: the stream is idle now
  for (smb_result* r : h->result_pool) destroy_result(r);
  h->d_pairs.release();
  h->d_items.release();
  h->d_acc.release();
  h->d_log.release();
  h->d_cta_busy.release();
  h->h_cta_busy.release();
  h->h_pairs.release();
  h->h_items.release();
  if (h->d_counters) cudaFree(h->d_counters);
  if (h->d_landed) cudaFree(h->d_landed);
  if (h->h_ticket_vals) cudaFreeHost(h->h_ticket_vals);
  if (h->lut_dev) cudaFree(h->lut_dev);
  if (h->pool) cudaFree(h->pool);
  if (h->kp_pool) cudaFree(h->kp_pool);
  h->d_pts.release();
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->stream_b) cudaStreamDestroy(h->stream_b);
  if (h->stream_up) cudaStreamDestroy(h->stream_up);
  for (auto& e : h->up_ev)
    if (e) cudaEventDestroy(e);
  if (h->ev_ext) cudaEventDestroy(h->ev_ext);
  if (h->ev_ext2) cudaEventDestroy(h->ev_ext2);
  for (auto& e : h->tail_ev)
    if (e) cudaEventDestroy(e);
  delete h;
}

int smb_set_options(smb_handle* h, const smb_options* opts) {
  if (!h) return SMB_EINVAL;
  return apply_options(h, opts);
}

int smb_put_image(smb_handle* h, uint32_t image_id, const uint8_t* desc, size_t n, size_t d) {
  if (!h) return SMB_EINVAL;
  SMB_CUDA(h, cudaSetDevice(h->device));
  int rc = put_image_impl(h, image_id, desc, n, d, cudaMemcpyHostToDevice);
  if (rc != SMB_OK) return rc;
  // the caller's buffer (Scanner-owned, valid only for the execute() call) is free again on return
  SMB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SMB_OK;
}

int smb_put_images(smb_handle* h, const uint32_t* image_ids, const uint8_t* const* descs, const size_t* ns,
                   size_t count, size_t d) {
  if (!h) return SMB_EINVAL;
  if (count && (!image_ids || !descs || !ns)) return fail(h, SMB_EINVAL, "null array argument");
  SMB_CUDA(h, cudaSetDevice(h->device));
  for (size_t k = 0; k < count; ++k) {
    int rc = put_image_impl(h, image_ids[k], descs[k], ns[k], d, cudaMemcpyHostToDevice);
    if (rc != SMB_OK) return rc;
  }
  SMB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SMB_OK;
}

int smb_put_images_async(smb_handle* h, const uint32_t* image_ids, const uint8_t* const* descs, const size_t* ns,
                         size_t count, size_t d) {
  if (!h) return SMB_EINVAL;
  if (count && (!image_ids || !descs || !ns)) return fail(h, SMB_EINVAL, "null array argument");
  SMB_CUDA(h, cudaSetDevice(h->device));
  if (h->up_issued - h->up_synced >= (uint64_t)smb_handle::kUpRing - 1)  // the event ring is about to wrap
    if (int rc = drain_uploads(h)) return rc;
  const uint64_t ticket = ++h->up_issued;
  h->up_fast[ticket % smb_handle::kUpRing] = 0;  // host upload over PCIe: worth a sub-batch of its own
  OpenTicket open(h, ticket);
  for (size_t k = 0; k < count; ++k) {
    int rc = put_image_impl(h, image_ids[k], descs[k], ns[k], d, cudaMemcpyHostToDevice, h->stream_up, ticket);
    if (rc != SMB_OK) return rc;
  }
  SMB_CUDA(h, open.finish());
  return SMB_OK;
}

int smb_put_image_device(smb_handle* h, uint32_t image_id, const void* dev_desc, size_t n, size_t d) {
  if (!h) return SMB_EINVAL;
  SMB_CUDA(h, cudaSetDevice(h->device));
  int rc = put_image_impl(h, image_id, dev_desc, n, d, cudaMemcpyDefault);
  if (rc != SMB_OK) return rc;
  SMB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SMB_OK;
}

int smb_put_images_device(smb_handle* h, const uint32_t* image_ids, const void* const* dev_descs, const size_t* ns,
                          size_t count, size_t d) {
  if (!h) return SMB_EINVAL;
  if (count && (!image_ids || !dev_descs || !ns)) return fail(h, SMB_EINVAL, "null array argument");
  SMB_CUDA(h, cudaSetDevice(h->device));
  for (size_t k = 0; k < count; ++k) {
    int rc = put_image_impl(h, image_ids[k], dev_descs[k], ns[k], d, cudaMemcpyDefault);
    if (rc != SMB_OK) return rc;
  }
  SMB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SMB_OK;
}

int smb_put_images_device_async(smb_handle* h, const uint32_t* image_ids, const void* const* dev_descs,
                                const size_t* ns, size_t count, size_t d, void* producer_stream) {
  if (!h) return SMB_EINVAL;
  if (count && (!image_ids || !dev_descs || !ns)) return fail(h, SMB_EINVAL, "null array argument");
  SMB_CUDA(h, cudaSetDevice(h->device));
  if (h->up_issued - h->up_synced >= (uint64_t)smb_handle::kUpRing - 1)  // the event ring is about to wrap
    if (int rc = drain_uploads(h)) return rc;
  const uint64_t ticket = ++h->up_issued;
  h->up_fast[ticket % smb_handle::kUpRing] = 1;  // device-to-device adoption: never splits a match call
  OpenTicket open(h, ticket);
  // the copies below must not start before the producer's queued work (e.g. the NCCL recv) has finished
  SMB_CUDA(h, cudaEventRecord(h->ev_ext, static_cast<cudaStream_t>(producer_stream)));
  SMB_CUDA(h, cudaStreamWaitEvent(h->stream_up, h->ev_ext, 0));
  for (size_t k = 0; k < count; ++k) {
    int rc = put_image_impl(h, image_ids[k], dev_descs[k], ns[k], d, cudaMemcpyDefault, h->stream_up, ticket);
    if (rc != SMB_OK) return rc;
  }
  SMB_CUDA(h, open.finish());
  return SMB_OK;
}

int smb_has_image(const smb_handle* h, uint32_t image_id) { return h && h->images.count(image_id) ? 1 : 0; }

int smb_evict_image(smb_handle* h, uint32_t image_id) {
  if (!h) return SMB_EINVAL;
  auto it = h->images.find(image_id);
  if (it == h->images.end()) return fail(h, SMB_EINVAL, "image %u is not cached", image_id);
  SMB_CUDA(h, cudaSetDevice(h->device));
  if (int rc = retire_rows(h, it->second)) return rc;  // no device synchronisation (see retire_rows)
  h->images.erase(it);
  ++h->layout_epoch;
  return SMB_OK;
}

int smb_clear_images(smb_handle* h) {
  if (!h) return SMB_EINVAL;
  SMB_CUDA(h, cudaSetDevice(h->device));
  if (h->inflight) return fail(h, SMB_EINVAL, "smb_clear_images while a match call is in flight");
  if (int rc = drain_uploads(h)) return rc;  // pending uploads still write pool rows
  h->images.clear();
  h->free_list.clear();
  if (h->pool_rows) h->free_list.emplace(0u, h->pool_rows);
  ++h->layout_epoch;
  return SMB_OK;
}

int smb_image_device_ptr(const smb_handle* h, uint32_t image_id, const void** dev_ptr, size_t* n) {
  if (!h || !dev_ptr || !n) return SMB_EINVAL;
  auto it = h->images.find(image_id);
  if (it == h->images.end()) return fail(const_cast<smb_handle*>(h), SMB_EINVAL, "image %u is not cached", image_id);
  *dev_ptr = it->second.n ? h->pool + (size_t)it->second.row0 * kDim : nullptr;
  *n = it->second.n;
  return SMB_OK;
}

// One match call = enqueue (everything the device needs, no host decision in between) + finish (one stream
// synchronisation, then the counters the device left in pinned memory say whether the attempt stands).
//
//   [plan fetch, only if the plan differs from the one already on the device]
//   per sub-batch: zero the accumulators -> score_tcgen05_kernel -> runner_up_kernel -> decide_kernel
//   copy of the 48-byte counter block
//
// all on the handle's main stream.  decide_kernel writes the matches and the per-pair {start, count} table straight
// into the result's pinned host memory (zero-copy), so when the stream is idle the result is complete: no staging
// buffer, no device-to-host copy of the matches, no "how many matches?" round trip.
//
// Sub-batches exist (a) to bound the accumulator footprint and (b) to start on pairs whose images have landed while
// later HOST uploads (smb_put_images_async, PCIe) are still in flight: a new sub-batch begins where a pair needs a
// newer, still pending host-upload ticket than all pairs before it, and each sub-batch's kernels wait on the
// device for their own ticket only.  A pending device-to-device adoption (smb_put_images_device_async: the NVLink
// halo, microseconds) never splits a call -- every extra score launch pays its own tail (~0.2 ms measured), more
// than such a wait costs.
static cudaError_t reserve_acc(smb_handle* h, size_t n);
static int enqueue_match(smb_handle* h, smb_result* res, bool use_log, size_t matches_want) {
  const uint64_t* keys = res->keys.data();
  const size_t npairs = res->npairs;
  const bool prof = h->opts.profile != 0;
  const bool cc = h->opts.cross_check != 0;

  poll_uploads(h);  // uploads that have landed since the last look need no waiting (and no sub-batch of their own)

  // ---- plan: metas, work items, sub-batches
  std::vector<Sub> subs;
  std::vector<PairMeta>& pm = h->plan_pairs;
  std::vector<WorkItem>& wi = h->plan_items;
  // Two accumulator regions ping-pong between consecutive sub-batches (the decide of k runs under the score of k+1),
  // so a sub-batch may use half the budget; a call of >= 64 pairs is cut into `waves` sub-batches of equal pair count
  // so that only the last one's runner-up / decide (1 / waves of the result delivery) is exposed.
  const size_t sub_budget = std::max<size_t>(h->acc_budget / 2, 1);
  const size_t wave_pairs = npairs >= 256 && h->waves > 1 ? (npairs + h->waves - 1) / h->waves : npairs;
  size_t out_cap = 0, max_acc = 0;
  uint64_t ops = 0;
  auto is_host_ticket = [&](uint64_t t) { return t > h->up_synced && !h->up_fast[t % smb_handle::kUpRing]; };
#ifdef SMB_TEST_ENGINES
  const bool kernel_waits = h->opts.engine == SMB_ENGINE_TCGEN05;
#else
  const bool kernel_waits = true;
#endif
  bool pending_host = false;
  for (uint64_t t = h->up_synced + 1; t <= h->up_issued; ++t) pending_host = pending_host || !h->up_fast[t % smb_handle::kUpRing];
  const bool reuse = h->plan_epoch == h->layout_epoch && !pending_host && h->plan_acc_budget == sub_budget + (size_t)h->waves &&
                     h->plan_keys.size() == 2 * npairs && h->dev_plan_pairs == npairs && h->dev_plan_items == wi.size() &&
                     std::memcmp(h->plan_keys.data(), keys, 2 * npairs * sizeof(uint64_t)) == 0;
  if (reuse) {
    subs = h->plan_subs;
    out_cap = h->plan_out_cap;
    max_acc = h->plan_max_acc;
    ops = h->plan_ops;
    // images refreshed in place since the plan was made (the halo): their tickets complete in order, so the first
    // sub-batch waiting for the newest pending one covers them all
    for (Sub& sb : subs) sb.wait_ticket = 0;
    if (h->up_issued > h->up_synced) subs[0].wait_ticket = h->up_issued;
  } else {
  h->plan_epoch = 0;
  pm.resize(npairs);
  wi.clear();
  // Planned order: with host uploads still in flight, pairs are taken in the order their images land (stable), so an
  // early pair listed after a late one does not wait for the late one's upload.  order[q] = caller index.
  std::vector<uint32_t>& order = h->plan_order;
  order.clear();
  // every pair's images, looked up once (consecutive pairs mostly share image 1: one hash lookup saved per pair)
  std::vector<const ImageEntry*>& ent = h->plan_entries;
  ent.resize(2 * npairs);
  {
    uint64_t last_key = ~0ull;
    const ImageEntry* last_entry = nullptr;
    for (size_t k = 0; k < 2 * npairs; ++k) {
      if (keys[k] != last_key || !last_entry) {
        auto it = h->images.find(keys[k]);
        if (it == h->images.end())
          return fail(h, SMB_EINVAL, "pair %zu names image %llu which is not cached", k / 2, (unsigned long long)keys[k]);
        last_key = keys[k];
        last_entry = &it->second;
      }
      ent[k] = last_entry;
    }
  }
  if (h->up_synced < h->up_issued) {
    std::vector<uint64_t>& tk = h->plan_ticket;
    tk.resize(npairs);
    bool sorted = true;
    for (size_t p = 0; p < npairs; ++p) {
      uint64_t t = 0;
      if (is_host_ticket(ent[2 * p]->up_seq)) t = ent[2 * p]->up_seq;
      if (is_host_ticket(ent[2 * p + 1]->up_seq)) t = std::max(t, ent[2 * p + 1]->up_seq);
      tk[p] = t;
      if (p && tk[p] < tk[p - 1]) sorted = false;
    }
    if (!sorted) {
      order.resize(npairs);
      for (size_t p = 0; p < npairs; ++p) order[p] = (uint32_t)p;
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return tk[x] < tk[y]; });
    }
  }
  {
    Sub cur{0, 0, 0, 0, 0, 0, 0};
    for (size_t p = 0; p < npairs; ++p) {
      const size_t src = order.empty() ? p : order[p];
      const ImageEntry &a = *ent[2 * src], &b = *ent[2 * src + 1];
      const uint64_t ticket = std::max(a.up_seq, b.up_seq);
      uint64_t host_ticket = 0;
      if (is_host_ticket(a.up_seq)) host_ticket = a.up_seq;
      if (is_host_ticket(b.up_seq)) host_ticket = std::max(host_ticket, b.up_seq);
      const size_t need = (size_t)a.n + b.n;
      // Pending HOST uploads are waited for inside the score kernel, item by item (WorkItem::wait_ticket); with the
      // test engine, which has no such wait, they split the call like the accumulator budget does.
      const bool newer_upload = !kernel_waits && host_ticket > cur.split_ticket;
      if (p > cur.first && (cur.acc + need > sub_budget || newer_upload || p - cur.first >= wave_pairs)) {
        cur.last = p;
        subs.push_back(cur);
        cur = Sub{p, 0, wi.size(), 0, 0, cur.split_ticket, cur.wait_ticket};
      }
      cur.split_ticket = std::max(cur.split_ticket, host_ticket);
      // stream-level wait (before the sub-batch's first kernel): device-to-device adoptions -- their producer may be a
      // kernel (NCCL recv) that the persistent score CTAs would starve -- and, without in-kernel waits, everything
      {
        uint64_t dev_ticket = 0;
        for (uint64_t t : {a.up_seq, b.up_seq})
          if (t > h->up_synced && (!kernel_waits || h->up_fast[t % smb_handle::kUpRing])) dev_ticket = std::max(dev_ticket, t);
        cur.wait_ticket = std::max(cur.wait_ticket, dev_ticket);
      }
      const uint32_t item_ticket = kernel_waits ? (uint32_t)host_ticket : 0u;
      (void)ticket;
      PairMeta& m = pm[p];
      m.a_row0 = a.row0;
      m.n1 = a.n;
      m.b_row0 = b.row0;
      m.n2 = b.n;
      m.acc_off = (uint32_t)cur.acc;
      m.out_slot = (uint32_t)src;  // decide_kernel fills the caller-order table directly
      cur.acc += need;
      // work items: one per 256-row strip; the strips of a pair stay adjacent so that concurrently running CTAs
      // stream the same image 2 out of L2
      if (a.n && b.n) {
        const uint32_t n_btiles = (b.n + kTileCols - 1) / kTileCols;
        for (uint32_t r = 0; r < a.n; r += kStripRows)
          wi.push_back(WorkItem{a.row0 + r, b.row0, n_btiles, a.n - r > (uint32_t)kMTile ? 2u : 1u, m.acc_off + r,
                                m.acc_off + a.n, (uint32_t)(p - cur.first), item_ticket});
      }
      out_cap += cc ? std::min(a.n, b.n) : a.n;
      ops += 2ull * a.n * b.n * kDim;
    }
    cur.last = npairs;
    subs.push_back(cur);
    for (size_t k = 0; k < subs.size(); ++k) {
      Sub& sb = subs[k];
      sb.items = (k + 1 < subs.size() ? subs[k + 1].item0 : wi.size()) - sb.item0;
      max_acc = std::max(max_acc, sb.acc);
      // ragged sets: largest column counts first inside each sub-batch (stable: a pair's strips stay adjacent); items
      // that wait for an upload stay in landing order, behind everything that can start at once
      WorkItem* it0 = wi.data() + sb.item0;
      bool uniform = true;
      for (size_t x = 1; x < sb.items && uniform; ++x) uniform = it0[x].n_btiles == it0[0].n_btiles;
      if (!uniform)
        std::stable_sort(it0, it0 + sb.items, [](const WorkItem& x, const WorkItem& y) {
          return x.wait_ticket != y.wait_ticket ? x.wait_ticket < y.wait_ticket : x.n_btiles > y.n_btiles;
        });
    }
  }
  if (out_cap > 0xFFFFFF00ull) return fail(h, SMB_EINVAL, "match capacity exceeds 2^32 in one call");
  {
    // The upload stream is in order: a host upload an item waits for inside the kernel may be queued BEHIND a
    // device-to-device adoption whose producer is a kernel (the NCCL recv of the halo).  That adoption must have
    // completed before the persistent score CTAs occupy every SM, or they could starve the very kernel they wait for.
    uint32_t max_item_ticket = 0;
    for (const WorkItem& w : wi) max_item_ticket = std::max(max_item_ticket, w.wait_ticket);
    uint64_t guard = 0;
    for (uint64_t t = h->up_synced + 1; t <= h->up_issued && t < max_item_ticket; ++t)
      if (h->up_fast[t % smb_handle::kUpRing]) guard = t;
    if (guard) subs[0].wait_ticket = std::max(subs[0].wait_ticket, guard);
  }
  if (!pending_host && order.empty()) {  // reusable as long as the layout and the pair list stay the same
    h->plan_keys.assign(keys, keys + 2 * npairs);
    h->plan_subs = subs;
    h->plan_out_cap = out_cap;
    h->plan_max_acc = max_acc;
    h->plan_ops = ops;
    h->plan_acc_budget = sub_budget + (size_t)h->waves;
    h->plan_epoch = h->layout_epoch;   // (the device copy of the plan is made below)
  }
  }  // !reuse
  res->worst_case = out_cap;
  const size_t n_items_total = wi.size();
  const size_t acc_region = (max_acc + 15) / 16 * 16;  // region 1 starts here (region 0 at 0)
  use_log = use_log && h->log_cap;
#ifdef SMB_TEST_ENGINES
  use_log = use_log && h->opts.engine == SMB_ENGINE_TCGEN05;
#endif
  res->used_log = use_log;

  // ---- result buffers (pinned, written by the device): sized from what earlier calls produced, worst case on retry
  if (matches_want == 0) {
    matches_want = h->matches_per_pair > 0.0 ? (size_t)(1.5 * h->matches_per_pair * (double)npairs) + 4096 : out_cap / 2 + 4096;
    if (h->result_cap_override) matches_want = h->result_cap_override;
  }
  matches_want = std::min(std::max<size_t>(matches_want, 1), std::max<size_t>(out_cap, 1));
  res->matches_limit = matches_want;
  if (!reserve_pinned(&res->matches, &res->matches_cap, matches_want) || !reserve_pinned(&res->pair_out, &res->pair_cap, npairs))
    return fail(h, SMB_ENOMEM, "pinned result allocation failed (%zu matches, %zu pairs)", matches_want, npairs);
  if (cudaSuccess != h->d_pairs.reserve(npairs) || cudaSuccess != h->d_items.reserve(std::max<size_t>(n_items_total, 1)) ||
      cudaSuccess != reserve_acc(h, std::max<size_t>((subs.size() > 1 ? 2 : 1) * acc_region, 1)) ||
      cudaSuccess != h->h_pairs.reserve(npairs) ||
      cudaSuccess != h->h_items.reserve(std::max<size_t>(n_items_total, 1)) ||
      (use_log && cudaSuccess != h->d_log.reserve(h->log_cap))) {
    cudaGetLastError();
    return fail(h, SMB_ENOMEM, "device/pinned scratch allocation failed (pairs=%zu acc=%zu)", npairs, acc_region);
  }
  while (h->chain_ev.size() < 2 * subs.size()) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(h, SMB_ECUDA, "cudaEventCreate failed");
    h->chain_ev.push_back(e);
  }
  if (prof) {
    while (res->ev.size() < 5 * subs.size() + 2) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return fail(h, SMB_ECUDA, "cudaEventCreate failed");
      res->ev.push_back(e);
    }
  }
  if (prof && (cudaSuccess != h->d_cta_busy.reserve(subs.size() * (size_t)h->num_sms) ||
               cudaSuccess != h->h_cta_busy.reserve(subs.size() * (size_t)h->num_sms))) {
    cudaGetLastError();
    return fail(h, SMB_ENOMEM, "profiling scratch allocation failed");
  }
  res->n_subs = subs.size();
  res->sub_has_items.assign(subs.size(), 0);
  res->sub_grid.assign(subs.size(), 0);
  res->ops = ops;
  res->launches = res->score_launches = res->plan_uploaded = 0;

  cudaStream_t st = h->stream, sb2 = h->stream_b;
#define SMB_CUDA_R(expr)                                                                         \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      cudaStreamSynchronize(st);                                                                 \
      cudaStreamSynchronize(sb2);                                                                \
      h->acc_zero_slots = 0; /* kernels may have dirtied accumulators no decide_kernel cleared */  \
      return fail(h, SMB_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                            \
  } while (0)

  // ---- plan upload, skipped when the device already holds exactly this plan (steady state: same pairs, same rows).
  // The device reads the pinned plan directly (UVA): a cudaMemcpyAsync would queue behind whatever descriptor uploads
  // are already in the host->device copy engine's FIFO and stall the score kernel it feeds.
  const bool same_plan = reuse || (h->dev_plan_pairs == npairs && h->dev_plan_items == n_items_total &&
                                   std::memcmp(h->h_pairs.p, pm.data(), npairs * sizeof(PairMeta)) == 0 &&
                                   (n_items_total == 0 ||
                                    std::memcmp(h->h_items.p, wi.data(), n_items_total * sizeof(WorkItem)) == 0));
  if (prof) SMB_CUDA_R(cudaEventRecord(res->ev[0], st));
  SMB_CUDA_R(cudaEventRecord(h->tail_ev[0], st));
  SMB_CUDA_R(cudaMemsetAsync(h->d_counters, 0, smb_handle::kNumCounters * sizeof(unsigned long long), st));
  if (!same_plan) {
    h->dev_plan_pairs = h->dev_plan_items = 0;  // invalid until the fetch below has been queued
    std::memcpy(h->h_pairs.p, pm.data(), npairs * sizeof(PairMeta));
    if (n_items_total) std::memcpy(h->h_items.p, wi.data(), n_items_total * sizeof(WorkItem));
    static_assert(sizeof(PairMeta) % 4 == 0 && sizeof(WorkItem) % 4 == 0, "plan structs are whole words");
    const size_t wp = npairs * sizeof(PairMeta) / 4, ww = n_items_total * sizeof(WorkItem) / 4;
    fetch_words_kernel<<<(unsigned)std::min<size_t>((wp + 255) / 256, 512), 256, 0, st>>>(
        reinterpret_cast<uint32_t*>(h->d_pairs.p), reinterpret_cast<const uint32_t*>(h->h_pairs.p), wp);
    if (ww)
      fetch_words_kernel<<<(unsigned)std::min<size_t>((ww + 255) / 256, 512), 256, 0, st>>>(
          reinterpret_cast<uint32_t*>(h->d_items.p), reinterpret_cast<const uint32_t*>(h->h_items.p), ww);
    SMB_CUDA_R(cudaGetLastError());
    res->launches += ww ? 2 : 1;
    res->plan_uploaded = 1;
    h->dev_plan_pairs = npairs;
    h->dev_plan_items = n_items_total;
  }
  if (n_items_total && !h->tmap_valid) return fail(h, SMB_ECUDA, "descriptor pool tensor map is not initialised");

  // ---- the sub-batches, software-pipelined over two streams:
  //   main stream   : score(0)   score(1)              score(2)              ...
  //   second stream :            runner-up(0) decide(0) runner-up(1) decide(1) ...
  // Consecutive sub-batches alternate between two accumulator regions and the two halves of the survivor log;
  // score(k) waits for decide(k - 2), which frees its region and log half.  The runner-up and decide CTAs are small
  // enough to be resident next to a score CTA (kernels.cuh), so the PCIe-bound delivery of sub-batch k's matches
  // hides under the scoring of k + 1 and only the last sub-batch's is exposed.
  const size_t log_half = subs.size() > 1 ? h->log_cap / 2 : h->log_cap;
  uint64_t waited = h->up_synced;
  for (size_t k = 0; k < subs.size(); ++k) {
    const Sub& sb = subs[k];
    const size_t r = k & 1;
    TopTwo* acc = h->d_acc.p + r * acc_region;
    const SurvivorLog slog{use_log ? h->d_log.p + r * log_half : nullptr, h->d_counters + (r ? 5 : 2), (unsigned long long)log_half};
    cudaEvent_t ev_scored = h->chain_ev[2 * k], ev_decided = h->chain_ev[2 * k + 1];
    if (k >= 2) SMB_CUDA_R(cudaStreamWaitEvent(st, h->chain_ev[2 * (k - 2) + 1], 0));  // region r and log half r are free
    // tickets complete in order on the upload stream: waiting for the newest one this sub-batch needs is enough
    if (sb.wait_ticket > waited) {
      SMB_CUDA_R(cudaStreamWaitEvent(st, h->up_ev[sb.wait_ticket % smb_handle::kUpRing], 0));
      waited = sb.wait_ticket;
    }
    if (r * acc_region + sb.acc > h->acc_zero_slots) {  // only what no earlier decide_kernel has left clean: the whole
      const size_t upto = r * acc_region + sb.acc;       // prefix [0, acc_zero_slots) of d_acc is zero between calls
      SMB_CUDA_R(cudaMemsetAsync(h->d_acc.p + h->acc_zero_slots, 0, (upto - h->acc_zero_slots) * sizeof(TopTwo), st));
      h->acc_zero_slots = upto;
    }
    if (prof) SMB_CUDA_R(cudaEventRecord(res->ev[2 + 5 * k], st));
    if (sb.items) {
      res->sub_has_items[k] = 1;
      unsigned long long* cand = prof ? h->d_counters + 1 : nullptr;
#ifdef SMB_TEST_ENGINES
      if (h->opts.engine == SMB_ENGINE_DP4A) {
        const unsigned grid = (unsigned)std::min<size_t>(sb.items, (size_t)h->num_sms * 4);
        score_dp4a_kernel<<<grid, kDp4aThreads, 0, st>>>(h->pool, h->d_items.p + sb.item0, (uint32_t)sb.items,
                                                         h->d_pairs.p + sb.first, acc, h->filter.min_score, cand);
      } else
#endif
      {
        if (use_log) SMB_CUDA_R(cudaMemsetAsync(slog.count, 0, sizeof(unsigned long long), st));
        const unsigned grid = (unsigned)std::min<size_t>(sb.items, (size_t)h->num_sms);  // persistent CTAs
        score_tcgen05_kernel<<<grid, kScoreThreads, kScoreSmemBytes, st>>>(h->tmap, h->d_items.p + sb.item0, (uint32_t)sb.items,
                                                                          h->d_pairs.p + sb.first, acc, slog,
                                                                          h->filter.min_score, cand, h->dbg_flags,
                                                                          prof ? h->d_cta_busy.p + k * (size_t)h->num_sms : nullptr,
                                                                          h->d_landed);
        res->sub_grid[k] = grid;
      }
      SMB_CUDA_R(cudaGetLastError());
      res->score_launches++;
      res->launches++;
    }
    if (prof) SMB_CUDA_R(cudaEventRecord(res->ev[3 + 5 * k], st));
    if (k + 1 == subs.size()) SMB_CUDA_R(cudaEventRecord(h->tail_ev[1], st));
    SMB_CUDA_R(cudaEventRecord(ev_scored, st));
    // ---- second stream
    SMB_CUDA_R(cudaStreamWaitEvent(sb2, ev_scored, 0));
    if (prof) SMB_CUDA_R(cudaEventRecord(res->ev[4 + 5 * k], sb2));
    if (sb.items && use_log) {
      runner_up_kernel<<<(unsigned)h->num_sms * 16, 128, 0, sb2>>>(slog.entries, slog.count, h->d_counters + 3, slog.capacity, acc);
      SMB_CUDA_R(cudaGetLastError());
      res->launches++;
    }
    if (prof) SMB_CUDA_R(cudaEventRecord(res->ev[5 + 5 * k], sb2));
    if (k + 1 < subs.size())  // runs underneath the next sub-batch's score kernel: the small CTA shape
      decide_kernel<kDecideThreadsSmall><<<(unsigned)(sb.last - sb.first), kDecideThreadsSmall, 0, sb2>>>(
          h->d_pairs.p + sb.first, acc, h->lut_dev, h->max_ratio_f, h->max_distance_f, cc ? 1 : 0,
          reinterpret_cast<uint2*>(res->matches), (unsigned long long)res->matches_limit, h->d_counters, h->d_counters + 4,
          res->pair_out);
    else                      // runs alone
      decide_kernel<kDecideThreadsLarge><<<(unsigned)(sb.last - sb.first), kDecideThreadsLarge, 0, sb2>>>(
          h->d_pairs.p + sb.first, acc, h->lut_dev, h->max_ratio_f, h->max_distance_f, cc ? 1 : 0,
          reinterpret_cast<uint2*>(res->matches), (unsigned long long)res->matches_limit, h->d_counters, h->d_counters + 4,
          res->pair_out);
    SMB_CUDA_R(cudaGetLastError());
    res->launches++;
    if (prof) SMB_CUDA_R(cudaEventRecord(res->ev[6 + 5 * k], sb2));
    SMB_CUDA_R(cudaEventRecord(ev_decided, sb2));
  }
  // the second stream has seen every score kernel and ran every decide: the counters (and profiling data) last
  SMB_CUDA_R(cudaMemcpyAsync(res->counters, h->d_counters, smb_handle::kNumCounters * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, sb2));
  if (prof)
    SMB_CUDA_R(cudaMemcpyAsync(h->h_cta_busy.p, h->d_cta_busy.p, subs.size() * (size_t)h->num_sms * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, sb2));
  if (prof) SMB_CUDA_R(cudaEventRecord(res->ev[1], sb2));
  SMB_CUDA_R(cudaEventRecord(h->tail_ev[2], sb2));
#undef SMB_CUDA_R
  return SMB_OK;
}

static cudaError_t reserve_acc(smb_handle* h, size_t n) {
  TopTwo* before = h->d_acc.p;
  cudaError_t e = h->d_acc.reserve(n);
  if (h->d_acc.p != before) h->acc_zero_slots = 0;  // a new allocation holds garbage
  return e;
}

static void flush_pending_free(smb_handle* h) {
  for (auto& f : h->pending_free) free_rows(h, f.first, f.second);
  h->pending_free.clear();
}

// Wait for an enqueued call and decide whether its results stand.  Two (rare) reasons to repeat it: the survivor log
// overflowed (adversarial inputs where almost every score survives: repeat with the returning two-stage insertion),
// or the matches did not fit the result buffer (repeat with the worst-case size).
static int finish_match(smb_handle* h, smb_result* res) {
  for (int attempt = 0;; ++attempt) {
    cudaError_t e = cudaStreamSynchronize(h->stream_b);  // the call's last work is on the second stream
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
      h->acc_zero_slots = 0;
      return fail(h, SMB_ECUDA, "match call failed on the device: %s", cudaGetErrorString(e));
    }
    const unsigned long long* c = res->counters;
    const bool log_over = res->used_log && c[3] != 0, out_over = c[4] != 0;
    if (!log_over && !out_over) break;
    if (attempt >= 2) return fail(h, SMB_ECUDA, "internal error: match call still overflows after two repeats");
    int rc = enqueue_match(h, res, res->used_log && !log_over, out_over ? res->worst_case : res->matches_limit);
    if (rc != SMB_OK) return rc;
  }
  res->total = (size_t)res->counters[0];
  if (h->waves_adaptive && h->waves == 1 && res->npairs >= 256) {
    float call_ms = 0.f, tail_ms = 0.f;
    if (cudaEventElapsedTime(&call_ms, h->tail_ev[0], h->tail_ev[2]) == cudaSuccess &&
        cudaEventElapsedTime(&tail_ms, h->tail_ev[1], h->tail_ev[2]) == cudaSuccess && tail_ms > 0.12f * call_ms)
      h->waves = 4;
    cudaGetLastError();
  }
  if (res->npairs) h->matches_per_pair = std::max(h->matches_per_pair, (double)res->total / (double)res->npairs);
  std::memset(&h->timing, 0, sizeof h->timing);
  h->timing.ops = res->ops;
  h->timing.score_launches = res->score_launches;
  h->timing.total_launches = res->launches;
  h->timing.sub_batches = (uint32_t)res->n_subs;
  h->timing.plan_uploaded = res->plan_uploaded;
  if (h->opts.profile && res->ev.size() >= 5 * res->n_subs + 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, res->ev[0], res->ev[1]) == cudaSuccess) h->timing.total_ms = ms;
    for (size_t k = 0; k < res->n_subs; ++k) {  // (runner-up / decide of sub-batch k overlap the score kernel of k + 1)
      if (cudaEventElapsedTime(&ms, res->ev[2 + 5 * k], res->ev[3 + 5 * k]) == cudaSuccess) h->timing.score_ms += ms;
      if (cudaEventElapsedTime(&ms, res->ev[4 + 5 * k], res->ev[5 + 5 * k]) == cudaSuccess) h->timing.runner_up_ms += ms;
      if (cudaEventElapsedTime(&ms, res->ev[5 + 5 * k], res->ev[6 + 5 * k]) == cudaSuccess) h->timing.decide_ms += ms;
    }
    cudaGetLastError();
    h->timing.candidates = res->counters[1];
    double sum_max = 0.0, sum_mean = 0.0;
    for (size_t k = 0; k < res->n_subs && k < res->sub_grid.size(); ++k) {
      const uint32_t g = res->sub_grid[k];
      if (!g) continue;
      const unsigned long long* b = h->h_cta_busy.p + k * (size_t)h->num_sms;
      unsigned long long mx = 0, tot = 0;
      for (uint32_t c = 0; c < g; ++c) { mx = std::max(mx, b[c]); tot += b[c]; }
      sum_max += (double)mx;
      sum_mean += (double)tot / g;
    }
    h->timing.cta_busy_max_over_mean = sum_mean > 0.0 ? (float)(sum_max / sum_mean) : 0.f;
  }
  return SMB_OK;
}

static int begin_keys(smb_handle* h, std::vector<uint64_t>&& keys, size_t npairs, smb_result** out) {
  *out = nullptr;
  if (h->inflight) return fail(h, SMB_EINVAL, "a match call is already in flight on this handle (smb_result_wait it first)");
  if (npairs > 0x7FFFFFFFull) return fail(h, SMB_EINVAL, "too many pairs in one call");
  SMB_CUDA(h, cudaSetDevice(h->device));
  smb_result* res = acquire_result(h);
  if (!res) return fail(h, SMB_ENOMEM, "out of host memory");
  res->keys = std::move(keys);
  res->npairs = npairs;
  res->total = 0;
  res->pending = false;
  if (npairs == 0) {
    *out = res;
    return SMB_OK;
  }
  int rc = enqueue_match(h, res, /*use_log=*/true, /*matches_want=*/0);
  if (rc != SMB_OK) {
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->stream_b);
    h->result_pool.push_back(res);
    return rc;
  }
  res->pending = true;
  res->verified = false;
  h->inflight = res;
  h->last_result = res;
  *out = res;
  return SMB_OK;
}

int smb_result_wait(smb_handle* h, smb_result* r) {
  if (!h || !r) return SMB_EINVAL;
  if (!r->pending) return SMB_OK;
  if (h->inflight != r) return fail(h, SMB_EINVAL, "result does not belong to the call in flight on this handle");
  SMB_CUDA(h, cudaSetDevice(h->device));
  const int rc = finish_match(h, r);
  r->pending = false;
  h->inflight = nullptr;
  flush_pending_free(h);
  if (rc != SMB_OK) r->npairs = 0;  // nothing valid to read
  return rc;
}

int smb_match_pairs_begin(smb_handle* h, const uint32_t* pairs, size_t npairs, smb_result** out) {
  if (!h) return SMB_EINVAL;
  if (!out) return fail(h, SMB_EINVAL, "out pointer is null");
  if (npairs && !pairs) return fail(h, SMB_EINVAL, "pairs pointer is null");
  std::vector<uint64_t> keys(2 * npairs);
  for (size_t k = 0; k < 2 * npairs; ++k) keys[k] = pairs[k];
  return begin_keys(h, std::move(keys), npairs, out);
}

int smb_match_pairs(smb_handle* h, const uint32_t* pairs, size_t npairs, smb_result** out) {
  int rc = smb_match_pairs_begin(h, pairs, npairs, out);
  if (rc != SMB_OK) return rc;
  rc = smb_result_wait(h, *out);
  if (rc != SMB_OK) {
    smb_result_release(h, *out);
    *out = nullptr;
  }
  return rc;
}

size_t smb_result_num_pairs(const smb_result* r) { return r && !r->pending ? r->npairs : 0; }

const smb_match* smb_result_matches(const smb_result* r, size_t i, size_t* count) {
  if (!r || r->pending || i >= r->npairs) {
    if (count) *count = 0;
    return nullptr;
  }
  if (count) *count = r->pair_out[i].count;
  return r->matches ? r->matches + r->pair_out[i].start : nullptr;
}

size_t smb_result_total_matches(const smb_result* r) { return r && !r->pending ? r->total : 0; }

void smb_result_release(smb_handle* h, smb_result* r) {
  if (!r) return;
  if (h) {
    if (r->pending) smb_result_wait(h, r);  // never recycle buffers the device may still write
    if (h->last_result == r) h->last_result = nullptr;
    r->npairs = 0;
    r->verified = false;
    h->result_pool.push_back(r);
  } else {
    destroy_result(r);
  }
}

int smb_match_descriptors(smb_handle* h, const uint8_t* desc1, size_t n1, const uint8_t* desc2, size_t n2,
                          smb_match* out, size_t capacity, size_t* count) {
  if (!h) return SMB_EINVAL;
  if (!count) return fail(h, SMB_EINVAL, "count pointer is null");
  *count = 0;
  SMB_CUDA(h, cudaSetDevice(h->device));
  const uint64_t k1 = 1ull << 32, k2 = (1ull << 32) + 1;  // private ids, cannot collide with uint32 image ids
  int rc = put_image_impl(h, k1, desc1, n1, kDim, cudaMemcpyHostToDevice);
  if (rc == SMB_OK) rc = put_image_impl(h, k2, desc2, n2, kDim, cudaMemcpyHostToDevice);
  smb_result* res = nullptr;
  if (rc == SMB_OK) {
    rc = begin_keys(h, std::vector<uint64_t>{k1, k2}, 1, &res);
    if (rc == SMB_OK) {
      rc = smb_result_wait(h, res);
      if (rc != SMB_OK) {
        smb_result_release(h, res);
        res = nullptr;
      }
    }
  }
  if (rc == SMB_OK) {
    size_t c = 0;
    const smb_match* m = smb_result_matches(res, 0, &c);
    if (c > capacity) {
      rc = fail(h, SMB_ECAPACITY, "output capacity %zu < %zu matches", capacity, c);
    } else {
      if (c) std::memcpy(out, m, c * sizeof(smb_match));
      *count = c;
    }
    smb_result_release(h, res);
  }
  for (uint64_t k : {k1, k2}) {  // the call above has completed: nothing reads these rows any more
    auto it = h->images.find(k);
    if (it != h->images.end()) {
      free_rows(h, it->second.row0, it->second.rows);
      h->images.erase(it);
      ++h->layout_epoch;
    }
  }
  return rc;
}

int smb_get_timing(const smb_handle* h, smb_timing* t) {
  if (!h || !t) return SMB_EINVAL;
  *t = h->timing;
  return SMB_OK;
}

void smb_default_tvg_options(smb_tvg_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);  // flags = 0: watermark detection on, one model
  o->min_num_inliers = 15;     // colmap.proto:41
  o->min_num_trials = 30;      // colmap.proto:32
  o->max_num_trials = 10000;   // colmap.proto:33
  o->max_error = 4.0;          // colmap.proto:26
  o->confidence = 0.999;       // colmap.proto:29
  o->min_inlier_ratio = 0.25;  // colmap.proto:37
  o->max_h_inlier_ratio = 0.8;
  o->seed = 0;
}

int smb_put_keypoints(smb_handle* h, uint32_t image_id, const float* xy, size_t n, size_t stride_bytes) {
  if (!h) return SMB_EINVAL;
  auto it = h->images.find(image_id);
  if (it == h->images.end()) return fail(h, SMB_EINVAL, "image %u is not cached", image_id);
  if (it->second.n != n) return fail(h, SMB_EINVAL, "image %u has %u descriptors, got %zu keypoints", image_id, it->second.n, n);
  if (n && (!xy || stride_bytes < sizeof(float2))) return fail(h, SMB_EINVAL, "bad keypoint array");
  SMB_CUDA(h, cudaSetDevice(h->device));
  if (n)
    SMB_CUDA(h, cudaMemcpy2DAsync(h->kp_pool + it->second.row0, sizeof(float2), xy, stride_bytes, sizeof(float2), n,
                                  cudaMemcpyHostToDevice, h->stream));
  SMB_CUDA(h, cudaStreamSynchronize(h->stream));  // the caller's buffer is free again on return
  it->second.has_kp = true;
  return SMB_OK;
}

int smb_result_verify(smb_handle* h, smb_result* r, const smb_tvg_options* opts) {
  if (!h || !r) return SMB_EINVAL;
  if (r->pending) return fail(h, SMB_EINVAL, "wait for the match call before verifying it");
  if (h->inflight) return fail(h, SMB_EINVAL, "another match call is in flight");
  if (h->last_result != r || h->dev_plan_pairs != r->npairs)
    return fail(h, SMB_EINVAL, "only the most recent completed match call of a handle can be verified");
  smb_tvg_options def;
  smb_default_tvg_options(&def);
  const smb_tvg_options& o = opts ? *opts : def;
  if (r->npairs == 0) {
    r->verified = true;
    return SMB_OK;
  }
  for (size_t k = 0; k < 2 * r->npairs; ++k) {
    auto it = h->images.find(r->keys[k]);
    if (it == h->images.end() || !it->second.has_kp)
      return fail(h, SMB_EINVAL, "image %llu has no keypoints (smb_put_keypoints)", (unsigned long long)r->keys[k]);
  }
  SMB_CUDA(h, cudaSetDevice(h->device));
  static_assert(sizeof(smb_tvg) == sizeof(tvg::Out), "smb_tvg mirrors tvg::Out");
  const size_t total = std::max<size_t>(r->total, 1);
  if (!reserve_pinned(&r->tvg, &r->tvg_cap, r->npairs) || !reserve_pinned(&r->inliers, &r->inliers_cap, total) ||
      cudaSuccess != h->d_pts.reserve(2 * total)) {  // correspondences of all matches + of the inliers
    cudaGetLastError();
    return fail(h, SMB_ENOMEM, "verification buffers (%zu pairs, %zu matches)", r->npairs, r->total);
  }
  tvg::Options to;
  to.min_num_inliers = o.min_num_inliers;
  to.min_num_trials = o.min_num_trials;
  to.max_num_trials = o.max_num_trials;
  to.max_error = o.max_error;
  to.confidence = o.confidence;
  to.min_inlier_ratio = o.min_inlier_ratio;
  to.max_H_inlier_ratio = o.max_h_inlier_ratio;
  to.watermark_min_inlier_ratio = 0.7;  // TwoViewGeometry::Options default, not exposed by the reference's proto
  to.detect_watermark = (o.flags & SMB_TVG_NO_WATERMARK) ? 0 : 1;
  to.pad_ = 0;
  to.seed = o.seed;
  if (!(o.flags & SMB_TVG_MULTIPLE_MODELS)) {
    tvg::verify_kernel<<<(unsigned)r->npairs, tvg::kThreads, 0, h->stream>>>(
        h->d_pairs.p, reinterpret_cast<const uint2*>(r->matches), r->pair_out, h->kp_pool, h->d_pts.p, h->d_pts.p + total, to,
        reinterpret_cast<tvg::Out*>(r->tvg), reinterpret_cast<uint2*>(r->inliers));
    SMB_CUDA(h, cudaGetLastError());
    SMB_CUDA(h, cudaStreamSynchronize(h->stream));
    r->verified = true;
    return SMB_OK;
  }
  // ---- TwoViewGeometry::EstimateMultiple: rounds of the same kernel over the matches no accepted model has explained yet.
  // The bookkeeping between rounds (a few thousand matches per pair) runs on the host; a round costs one launch + one sync.
  if (!reserve_pinned(&r->rem_matches, &r->rem_matches_cap, total) || !reserve_pinned(&r->rem_po, &r->rem_po_cap, r->npairs) ||
      !reserve_pinned(&r->round_tvg, &r->round_tvg_cap, r->npairs) || !reserve_pinned(&r->round_inl, &r->round_inl_cap, total))
    return fail(h, SMB_ENOMEM, "verification buffers (%zu pairs, %zu matches)", r->npairs, r->total);
  std::memcpy(r->rem_matches, r->matches, r->total * sizeof(smb_match));
  std::memcpy(r->rem_po, r->pair_out, r->npairs * sizeof(PairOut));
  std::vector<uint8_t> done(r->npairs, 0);
  std::vector<uint32_t> models(r->npairs, 0), appended(r->npairs, 0);
  std::vector<int32_t> trials_f(r->npairs, 0), trials_h(r->npairs, 0);
  for (size_t i = 0; i < r->npairs; ++i) {
    smb_tvg& t = r->tvg[i];
    std::memset(&t, 0, sizeof t);
    t.config = tvg::kDegenerate;
    t.inlier_start = r->pair_out[i].start;
  }
  for (uint32_t round = 0;; ++round) {
    to.seed = o.seed + 7919ull * round;
    tvg::verify_kernel<<<(unsigned)r->npairs, tvg::kThreads, 0, h->stream>>>(
        h->d_pairs.p, reinterpret_cast<const uint2*>(r->rem_matches), r->rem_po, h->kp_pool, h->d_pts.p, h->d_pts.p + total, to,
        reinterpret_cast<tvg::Out*>(r->round_tvg), reinterpret_cast<uint2*>(r->round_inl));
    SMB_CUDA(h, cudaGetLastError());
    SMB_CUDA(h, cudaStreamSynchronize(h->stream));
    bool any_active = false;
    for (size_t i = 0; i < r->npairs; ++i) {
      if (done[i]) continue;
      const smb_tvg& rt = r->round_tvg[i];
      trials_f[i] += rt.trials_f;
      trials_h[i] += rt.trials_h;
      if (rt.config == tvg::kDegenerate) {
        done[i] = 1;
        r->rem_po[i].count = 0;  // later rounds return at once for this pair
        continue;
      }
      const smb_match* inl = r->round_inl + rt.inlier_start;
      const uint32_t n_inl = rt.inlier_count;
      if (rt.config != tvg::kWatermark) {  // multiple_ignore_watermark = true: a watermark's inliers are removed, not kept
        if (models[i] == 0) {
          r->tvg[i] = rt;  // a single accepted model is reported as it is
          r->tvg[i].inlier_start = r->pair_out[i].start;
        }
        std::memcpy(r->inliers + r->pair_out[i].start + appended[i], inl, n_inl * sizeof(smb_match));
        appended[i] += n_inl;
        ++models[i];
      }
      // ExtractOutlierMatches: the round's inliers are an ordered subsequence of the remaining matches
      smb_match* rem = r->rem_matches + r->rem_po[i].start;
      uint32_t keep = 0, q = 0;
      for (uint32_t x = 0; x < r->rem_po[i].count; ++x) {
        if (q < n_inl && rem[x].idx1 == inl[q].idx1 && rem[x].idx2 == inl[q].idx2) ++q;
        else rem[keep++] = rem[x];
      }
      r->rem_po[i].count = keep;
      if (n_inl == 0 || (int32_t)keep < o.min_num_inliers) {  // the next Estimate would be DEGENERATE
        done[i] = 1;
        r->rem_po[i].count = 0;
      } else {
        any_active = true;
      }
    }
    if (!any_active) break;
  }
  for (size_t i = 0; i < r->npairs; ++i) {
    smb_tvg& t = r->tvg[i];
    t.inlier_count = appended[i];
    t.trials_f = trials_f[i];
    t.trials_h = trials_h[i];
    if (models[i] > 1) {  // COLMAP: config = MULTIPLE, inlier_matches = all models' inliers, everything else default
      t.config = tvg::kMultiple;
      t.num_inliers_f = (int32_t)appended[i];
      t.num_inliers_h = 0;
      std::memset(t.F, 0, sizeof t.F);
      std::memset(t.H, 0, sizeof t.H);
    }
  }
  r->verified = true;
  return SMB_OK;
}

int smb_result_tvg(const smb_result* r, size_t i, smb_tvg* out) {
  if (!r || !out || !r->verified || i >= r->npairs) return SMB_EINVAL;
  *out = r->tvg[i];
  return SMB_OK;
}

const smb_match* smb_result_inliers(const smb_result* r, size_t i, size_t* count) {
  if (!r || !r->verified || i >= r->npairs) {
    if (count) *count = 0;
    return nullptr;
  }
  if (count) *count = r->tvg[i].inlier_count;
  return r->inliers + r->tvg[i].inlier_start;
}

int smb_wait_stream(smb_handle* h, void* stream) {
  if (!h) return SMB_EINVAL;
  SMB_CUDA(h, cudaSetDevice(h->device));
  SMB_CUDA(h, cudaEventRecord(h->ev_ext2, static_cast<cudaStream_t>(stream)));
  SMB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_ext2, 0));
  return SMB_OK;
}

int smb_alloc_pinned(size_t bytes, void** out) {
  if (!out) return SMB_EINVAL;
  *out = nullptr;
  if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return fail(nullptr, SMB_ENOMEM, "cudaMallocHost of %zu bytes failed", bytes);
  }
  return SMB_OK;
}

void smb_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

int smb_get_filter(const smb_handle* h, int32_t* min_score, int32_t* min_best) {
  if (!h) return SMB_EINVAL;
  if (min_score) *min_score = h->filter.min_score;
  if (min_best) *min_best = h->filter.min_best;
  return SMB_OK;
}

void* smb_stream(const smb_handle* h) { return h ? (void*)h->stream : nullptr; }

int smb_stream_wait_uploads(smb_handle* h, void* stream) {
  if (!h) return SMB_EINVAL;
  SMB_CUDA(h, cudaSetDevice(h->device));
  if (h->up_issued > h->up_synced)  // tickets complete in order: the newest one covers all
    SMB_CUDA(h, cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->up_ev[h->up_issued % smb_handle::kUpRing], 0));
  return SMB_OK;
}

int smb_synchronize(smb_handle* h) {
  if (!h) return SMB_EINVAL;
  SMB_CUDA(h, cudaSetDevice(h->device));
  SMB_CUDA(h, cudaStreamSynchronize(h->stream));
  SMB_CUDA(h, cudaStreamSynchronize(h->stream_b));
  return drain_uploads(h);
}

}  // extern "C"
