// Drop-in replacement for the Scanner op of
//   /root/reference/integration/op_cpp/sequential_matching.cc
// Same registered op name (SequentialMatchingCPU), same input / output columns, same kernel-args protobuf,
// same serialized rows (io.cc) -- so integration/feature_matching.py and the downstream incremental_mapping
// stage (incremental_mapping.cc:245-260) consume it unchanged.  What changes: the call
//   colmap::MatchSiftFeaturesCPU(sift_options_, descriptors1, descriptors2, &featureMatches)   (:154)
// is served by the B200 library behind include/smb.h, for ALL pairs of ALL rows of the batch in one call,
// with descriptors uploaded once per image (the reference re-deserialises every stencil entry per row and
// only looks at batch item 0, :106-108).  Per execute(): new images are staged into pinned memory by a few helper
// threads and uploaded asynchronously in chunks, the match call is begun, the pair-id rows are written while the
// GPU works, then the matches (which the device wrote into pinned host memory) are verified and serialised.
// All kernel instances of a process share one GPU context per device.
//
// Two-view geometry verification (:84-101, :157-178): with -DSMB_WITH_COLMAP the unchanged CPU
// colmap::TwoViewGeometry::Estimate[Multiple] runs on the GPU's matches; without COLMAP (this image) it runs on the
// GPU as well (smb_result_verify: the uncalibrated F / H LORANSAC path Estimate takes with the reference's dummy
// cameras -- a statistical contract, not bit parity); SMB_OP_VERIFY=none emits the raw matches.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>

#include "scanner/api/kernel.h"
#include "scanner/api/op.h"
#include "scanner/util/common.h"
#include "scanner/util/memory.h"

#include "../../include/smb.h"
#include "proto_lite.h"
#include "wire.h"

#ifdef SMB_WITH_COLMAP
#include <colmap/base/camera.h>
#include <colmap/estimators/two_view_geometry.h>
#include <colmap/feature/utils.h>
#endif

namespace {

// glog-style abort-on-failure, the reference's only error convention (io.cc:392-404)
#define SMB_CHECK(cond, ...)                                   \
  do {                                                         \
    if (!(cond)) {                                             \
      std::fprintf(stderr, "[SequentialMatchingCPU/B200] CHECK failed: %s: ", #cond); \
      std::fprintf(stderr, __VA_ARGS__);                       \
      std::fprintf(stderr, "\n");                              \
      std::abort();                                            \
    }                                                          \
  } while (0)

using smb_wire::FeatureMatch;
using smb_wire::TwoViewGeometry;

// A few helper threads that copy Scanner's pageable descriptor rows into the pinned staging buffer.  One thread
// copies ~10 GB/s, i.e. ~100 us per 8192-descriptor image -- more than the GPU needs to match that image against
// its overlap-1 neighbours (~60 us), so the op would be host-bound without them.
class CopyPool {
 public:
  struct Job { uint8_t* dst; const uint8_t* src; size_t n; };
  explicit CopyPool(int helpers) {
    for (int k = 0; k < helpers; ++k) threads_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  // Copies all jobs (split into 256 KiB pieces); returns when every byte has been copied.
  void run(const std::vector<Job>& jobs) {
    pieces_.clear();
    constexpr size_t kPiece = 256 << 10;
    for (const Job& j : jobs)
      for (size_t off = 0; off < j.n; off += kPiece) pieces_.push_back(Job{j.dst + off, j.src + off, std::min(kPiece, j.n - off)});
    if (pieces_.empty()) return;
    next_.store(0);
    done_.store(0);
    if (!threads_.empty() && pieces_.size() > 1) {
      {
        std::lock_guard<std::mutex> g(mu_);
        ++generation_;
      }
      cv_.notify_all();
    }
    work();
    while (done_.load(std::memory_order_acquire) < pieces_.size()) std::this_thread::yield();
  }

 private:
  void work() {
    for (;;) {
      const size_t k = next_.fetch_add(1);
      if (k >= pieces_.size()) return;
      std::memcpy(pieces_[k].dst, pieces_[k].src, pieces_[k].n);
      done_.fetch_add(1, std::memory_order_release);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      work();
    }
  }
  std::vector<std::thread> threads_;
  std::vector<Job> pieces_;
  std::atomic<size_t> next_{0}, done_{0};
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

// One GPU context per process and device, shared by every kernel instance on it.  Scanner creates one CPU kernel
// instance per pipeline instance (typically one per core): if each of them owned a matcher they would each
// allocate a descriptor pool, a survivor log and accumulators, each launch persistent kernels that fill every SM,
// and all pile onto device 0.  Here the instances take turns on one handle (their GPU sections are serialised
// by a mutex -- the kernels fill the whole GPU anyway), and an image another instance has already uploaded is a
// cache hit.  Device: $SMB_DEVICE, else siftArgs.gpu_index when it is not "-1", else 0.
struct SharedGpu {
  std::mutex mu;
  smb_handle* h = nullptr;
  int refs = 0;
  smb_options opts{};                 // what the handle is currently set to
  struct Cached { uint64_t n, sig, last_use; bool has_kp; };
  std::unordered_map<uint32_t, Cached> cached;
  uint64_t tick = 0;
  uint8_t* stage = nullptr;           // pinned staging buffer for this execute()'s new images
  size_t stage_cap = 0;
  std::unique_ptr<CopyPool> pool;
};

std::mutex g_registry_mu;
std::map<int, std::unique_ptr<SharedGpu>>& registry() {
  static std::map<int, std::unique_ptr<SharedGpu>> r;
  return r;
}

SharedGpu* acquire_gpu(int device, const smb_options& o) {
  std::lock_guard<std::mutex> g(g_registry_mu);
  auto& slot = registry()[device];
  if (!slot) slot.reset(new SharedGpu());
  SharedGpu* s = slot.get();
  std::lock_guard<std::mutex> g2(s->mu);
  if (!s->h) {
    const int rc = smb_create(device, &o, &s->h);
    SMB_CHECK(rc == SMB_OK, "smb_create: %s", smb_last_error(nullptr));
    s->opts = o;
    int helpers = 3;
    if (const char* e = std::getenv("SMB_OP_COPY_THREADS")) helpers = std::max(0, std::atoi(e) - 1);
    s->pool.reset(new CopyPool(helpers));
  }
  ++s->refs;
  return s;
}

void release_gpu(SharedGpu* s) {
  std::lock_guard<std::mutex> g(g_registry_mu);
  std::lock_guard<std::mutex> g2(s->mu);
  if (--s->refs == 0) {
    s->pool.reset();
    smb_destroy(s->h);
    s->h = nullptr;
    smb_free_pinned(s->stage);
    s->stage = nullptr;
    s->stage_cap = 0;
    s->cached.clear();
  }
}

// Cheap content signature of a descriptor matrix: its first and last 256 bytes plus 62 evenly spaced 64-byte
// samples (~4.5 KB of a 1 MiB image, < 1 us).  Together with the row count it tells a cached image from another
// image that merely carries the same id (ids restart at 0 in every prepare_image kernel instance,
// prepare_image.cc:12, so they collide across tables and jobs); hashing every byte of every stencil entry of
// every row would cost more host time than the GPU needs for the matching.
uint64_t signature(const uint8_t* p, size_t bytes) {
  uint64_t hsh = 0x9E3779B97F4A7C15ull ^ bytes;
  auto mix = [&](const uint8_t* q, size_t n) {
    for (size_t k = 0; k + 8 <= n; k += 8) {
      uint64_t w;
      std::memcpy(&w, q + k, 8);
      hsh = (hsh ^ w) * 0xFF51AFD7ED558CCDull;
      hsh ^= hsh >> 29;
    }
  };
  if (bytes <= 8192) {
    mix(p, bytes);
    for (size_t k = bytes & ~size_t(7); k < bytes; ++k) hsh = (hsh ^ p[k]) * 0x100000001B3ull;
    return hsh;
  }
  mix(p, 256);
  mix(p + bytes - 256, 256);
  const size_t step = (bytes - 512) / 62;
  for (size_t k = 0; k < 62; ++k) mix(p + 256 + k * step, 64);
  return hsh;
}

class SequentialMatchingB200Kernel : public scanner::StenciledBatchedKernel, public scanner::VideoKernel {
 public:
  explicit SequentialMatchingB200Kernel(const scanner::KernelConfig& config) : scanner::StenciledBatchedKernel(config) {
    // sequential_matching.cc:36-76: empty args == proto2 defaults (what feature_matching.py sends)
    smb_proto::parse(config.args.data(), config.args.size(), args_);
    smb_default_options(&opts_);
    opts_.max_ratio = args_.siftargs.max_ratio;
    opts_.max_distance = args_.siftargs.max_distance;
    opts_.cross_check = args_.siftargs.cross_check ? 1 : 0;
    opts_.max_num_matches = args_.siftargs.max_num_matches;
    int device = 0;
    if (const char* e = std::getenv("SMB_DEVICE")) device = std::atoi(e);
    else if (args_.siftargs.gpu_index != "-1") device = std::atoi(args_.siftargs.gpu_index.c_str());
    verbose_ = std::getenv("SMB_OP_VERBOSE") != nullptr;  // the reference printf's per pair (:130-134,150,171)
#ifndef SMB_WITH_COLMAP
    // Without COLMAP the geometric verification runs on the GPU (smb_result_verify: the uncalibrated F / H LORANSAC
    // path TwoViewGeometry::Estimate takes with the reference's dummy cameras; statistical contract).
    // SMB_OP_VERIFY=none emits the raw matches instead (config UNDEFINED) -- what the matcher-parity tests compare.
    const char* vm = std::getenv("SMB_OP_VERIFY");
    gpu_verify_ = !(vm && std::string(vm) == "none");
#endif
    smb_default_tvg_options(&tvg_opts_);
    tvg_opts_.min_num_inliers = args_.siftargs.min_num_inliers;        // sequential_matching.cc:63-75
    tvg_opts_.max_error = args_.siftargs.max_error;
    tvg_opts_.confidence = args_.siftargs.confidence;
    tvg_opts_.min_num_trials = args_.siftargs.min_num_trials;
    tvg_opts_.max_num_trials = args_.siftargs.max_num_trials;
    tvg_opts_.min_inlier_ratio = args_.siftargs.min_inlier_ratio;
    tvg_opts_.flags = args_.siftargs.multiple_models ? SMB_TVG_MULTIPLE_MODELS : 0;  // sequential_matching.cc:94-96
    gpu_ = acquire_gpu(device, opts_);
  }
  ~SequentialMatchingB200Kernel() override { release_gpu(gpu_); }

  // Whatever was cached belongs to rows Scanner will not continue: drop it (the reference is stateless, it
  // re-deserialises every stencil entry of every row, sequential_matching.cc:115-122).
  void reset() override { drop_cache(); }
  void new_stream(const std::vector<scanner::u8>&) override { drop_cache(); }

  void execute(const scanner::StenciledBatchedElements& input_cols, scanner::BatchedElements& output_cols) override {
    SMB_CHECK(input_cols.size() >= 3 && output_cols.size() >= 2, "expected 3 input and 2 output columns");
    const size_t batch = input_cols[0].size();
    // ---- 1. decode every row's stencil; collect the distinct images of the batch
    struct Row { std::vector<uint32_t> ids; std::vector<uint32_t> partners; std::vector<size_t> partner_stencil; };
    struct Img { uint32_t id; smb_wire::DescriptorView d; const scanner::Element* kp; };
    std::vector<Row> rows(batch);
    std::vector<Img> used;
    std::unordered_map<uint32_t, size_t> used_pos;
    std::vector<uint32_t> pairs;
    for (size_t b = 0; b < batch; ++b) {
      const scanner::Elements& id_st = input_cols[0][b];
      const scanner::Elements& desc_st = input_cols[2][b];
      SMB_CHECK(id_st.size() == desc_st.size() && id_st.size() == input_cols[1][b].size(), "stencil sizes differ");
      Row& r = rows[b];
      for (size_t s = 0; s < id_st.size(); ++s) {
        const uint32_t id = smb_wire::read_image_id(id_st[s].buffer, id_st[s].size);
        r.ids.push_back(id);
        if (used_pos.emplace(id, used.size()).second)
          used.push_back(Img{id, smb_wire::view_descriptors(desc_st[s].buffer, desc_st[s].size), &input_cols[1][b][s]});
      }
      // sequential_matching.cc:139-146: anchor = stencil[0]; partners in stencil order, skipping the anchor id
      // and ids already seen (absorbs Scanner's REPEAT_EDGE halo at the table tail)
      for (size_t s = 1; s < r.ids.size(); ++s) {
        const uint32_t id2 = r.ids[s];
        if (id2 == r.ids[0] || std::count(r.partners.begin(), r.partners.end(), id2) > 0) continue;
        r.partners.push_back(id2);
        r.partner_stencil.push_back(s);
        pairs.push_back(r.ids[0]);
        pairs.push_back(id2);
      }
    }

    std::lock_guard<std::mutex> lock(gpu_->mu);
    SharedGpu& g = *gpu_;
    smb_handle* h = g.h;
    if (std::memcmp(&g.opts, &opts_, sizeof opts_) != 0) {  // another instance runs with other kernel args
      SMB_CHECK(smb_set_options(h, &opts_) == SMB_OK, "%s", smb_last_error(h));
      g.opts = opts_;
    }
    ++g.tick;
    // ---- 2. descriptor cache: an image is uploaded once and reused by every row (and instance) that names it, as
    // long as the row still carries the bytes that were uploaded (row count + content signature)
    std::vector<size_t> fresh;
    size_t fresh_bytes = 0;
    for (size_t k = 0; k < used.size(); ++k) {
      const Img& im = used[k];
      const uint64_t sig = signature(im.d.data, im.d.rows * SMB_DESC_DIM);
      auto it = g.cached.find(im.id);
      if (it != g.cached.end() && it->second.n == im.d.rows && it->second.sig == sig && smb_has_image(h, im.id)) {
        it->second.last_use = g.tick;
        continue;
      }
      g.cached[im.id] = SharedGpu::Cached{im.d.rows, sig, g.tick, false};
      fresh.push_back(k);
      fresh_bytes += im.d.rows * SMB_DESC_DIM;
    }
    // ---- 3. new images: pageable rows -> pinned staging (helper threads) -> asynchronous uploads in up to three
    // growing chunks, so the DMA of chunk k runs under the staging copy of chunk k+1 and the match call can start
    // on the pairs whose images have landed (it waits on the device, per sub-batch, for its own upload ticket)
    if (!fresh.empty()) {
      SMB_CHECK(smb_synchronize(h) == SMB_OK, "%s", smb_last_error(h));  // the previous execute()'s DMAs read the staging buffer
      if (fresh_bytes > g.stage_cap) {
        smb_free_pinned(g.stage);
        g.stage_cap = std::max(fresh_bytes + fresh_bytes / 4, (size_t)48 << 20);
        void* p = nullptr;
        SMB_CHECK(smb_alloc_pinned(g.stage_cap, &p) == SMB_OK, "pinned staging buffer of %zu bytes", g.stage_cap);
        g.stage = static_cast<uint8_t*>(p);
      }
      const size_t nchunks = fresh.size() >= 12 ? 3 : fresh.size() >= 4 ? 2 : 1;
      size_t off = 0, lo = 0;
      for (size_t c = 0; c < nchunks; ++c) {
        const size_t hi = c + 1 == nchunks ? fresh.size() : std::max(lo + 1, fresh.size() * (c + 1) * (c + 2) / (nchunks * (nchunks + 1)));
        std::vector<CopyPool::Job> jobs;
        std::vector<uint32_t> ids;
        std::vector<const uint8_t*> ptrs;
        std::vector<size_t> ns;
        for (size_t k = lo; k < hi; ++k) {
          const Img& im = used[fresh[k]];
          const size_t nb = im.d.rows * SMB_DESC_DIM;
          jobs.push_back(CopyPool::Job{g.stage + off, im.d.data, nb});
          ids.push_back(im.id);
          ptrs.push_back(g.stage + off);
          ns.push_back(im.d.rows);
          off += nb;
        }
        g.pool->run(jobs);
        const int rc = smb_put_images_async(h, ids.data(), ptrs.data(), ns.data(), ids.size(), SMB_DESC_DIM);
        SMB_CHECK(rc == SMB_OK, "smb_put_images_async: %s", smb_last_error(h));
        lo = hi;
      }
    }
    // ---- 4. every pair of the batch in one GPU call; the pair-id rows are serialised while it runs
    smb_result* res = nullptr;
    int rc = smb_match_pairs_begin(h, pairs.data(), pairs.size() / 2, &res);
    SMB_CHECK(rc == SMB_OK, "smb_match_pairs_begin: %s", smb_last_error(h));
    for (size_t b = 0; b < batch; ++b) {
      const Row& r = rows[b];
      const size_t n0 = smb_wire::pair_ids_bytes(r.partners.size());
      scanner::u8* b0 = scanner::new_buffer(scanner::CPU_DEVICE, n0);
      smb_wire::write_pair_ids(b0, r.partners);
      scanner::insert_element(output_cols[0], b0, n0);                  // io.cc:151-176
    }
    rc = smb_result_wait(h, res);
    SMB_CHECK(rc == SMB_OK, "smb_result_wait: %s", smb_last_error(h));
    if (gpu_verify_) {
      for (const Img& im : used) {
        SharedGpu::Cached& c = g.cached[im.id];
        if (c.has_kp) continue;
        const smb_wire::KeypointView kv = smb_wire::view_keypoints(im.kp->buffer, im.kp->size);
        SMB_CHECK(kv.n == im.d.rows, "image %u: %zu keypoints for %zu descriptors", im.id, kv.n, im.d.rows);
        rc = smb_put_keypoints(h, im.id, reinterpret_cast<const float*>(kv.data), kv.n, sizeof(smb_wire::FeatureKeypoint));
        SMB_CHECK(rc == SMB_OK, "smb_put_keypoints: %s", smb_last_error(h));
        c.has_kp = true;
      }
      tvg_opts_.seed = g.tick;
      rc = smb_result_verify(h, res, &tvg_opts_);
      SMB_CHECK(rc == SMB_OK, "smb_result_verify: %s", smb_last_error(h));
    }
    // ---- 5. verify + serialise, one output row per batch item (the reference emits one, for item 0 only)
    size_t p = 0;
    for (size_t b = 0; b < batch; ++b) {
      const Row& r = rows[b];
      std::vector<TwoViewGeometry> tvgs;
      tvgs.reserve(r.partners.size());
      for (size_t k = 0; k < r.partners.size(); ++k, ++p) {
        size_t m = 0;
        const smb_match* matches = smb_result_matches(res, p, &m);
        TwoViewGeometry tvg = gpu_verify_ ? from_gpu(res, p)
                                          : verify(matches, m, input_cols[1][b][0], input_cols[1][b][r.partner_stencil[k]]);
        if (verbose_) std::printf("View geometry for #%u and #%u has %zu inliers\n", r.ids[0], r.partners[k], tvg.inlier_matches.size());
        // sequential_matching.cc:173-178: too few inliers -> default-constructed TwoViewGeometry
        if (tvg.inlier_matches.size() < static_cast<size_t>(args_.siftargs.min_num_inliers)) tvg = TwoViewGeometry();
        tvgs.push_back(std::move(tvg));
      }
      const size_t n1 = smb_wire::tvg_list_bytes(tvgs);
      scanner::u8* b1 = scanner::new_buffer(scanner::CPU_DEVICE, n1);
      smb_wire::write_tvg_list(b1, tvgs);
      scanner::insert_element(output_cols[1], b1, n1);                  // io.cc:256-304
    }
    smb_result_release(h, res);
    // ---- 6. lazy eviction: images no execute() has named for two calls (the window has moved past them); no
    // device synchronisation is involved (smb_evict_image only edits the pool's free list)
    for (auto it = g.cached.begin(); it != g.cached.end();) {
      if (it->second.last_use + 2 <= g.tick) {
        if (smb_has_image(h, it->first)) SMB_CHECK(smb_evict_image(h, it->first) == SMB_OK, "%s", smb_last_error(h));
        it = g.cached.erase(it);
      } else ++it;
    }
  }

 private:
  void drop_cache() {
    std::lock_guard<std::mutex> lock(gpu_->mu);
    SMB_CHECK(smb_clear_images(gpu_->h) == SMB_OK, "%s", smb_last_error(gpu_->h));
    gpu_->cached.clear();
  }

  // The GPU verifier's verdict as a colmap::TwoViewGeometry row: config, F and H (Eigen stores them column-major,
  // io.cc:283-293), inlier_matches; E / qvec / tvec / tri_angle stay zero as EstimateUncalibrated leaves them.
  static TwoViewGeometry from_gpu(const smb_result* res, size_t p) {
    TwoViewGeometry out;
    smb_tvg t;
    if (smb_result_tvg(res, p, &t) != SMB_OK) return out;
    out.config = t.config;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        out.F[3 * c + r] = t.F[3 * r + c];
        out.H[3 * c + r] = t.H[3 * r + c];
      }
    size_t n = 0;
    const smb_match* in = smb_result_inliers(res, p, &n);
    out.inlier_matches.assign(reinterpret_cast<const FeatureMatch*>(in), reinterpret_cast<const FeatureMatch*>(in) + n);
    return out;
  }

  // converted from colmap::TwoViewGeometryVerifier::Run via sequential_matching.cc:84-101
  TwoViewGeometry verify(const smb_match* matches, size_t m, const scanner::Element& kp1, const scanner::Element& kp2) {
    TwoViewGeometry out;
#ifdef SMB_WITH_COLMAP
    const smb_wire::KeypointView k1 = smb_wire::view_keypoints(kp1.buffer, kp1.size), k2 = smb_wire::view_keypoints(kp2.buffer, kp2.size);
    colmap::FeatureKeypoints c1(k1.n), c2(k2.n);
    std::memcpy(c1.data(), k1.data, k1.n * sizeof(colmap::FeatureKeypoint));
    std::memcpy(c2.data(), k2.data, k2.n * sizeof(colmap::FeatureKeypoint));
    colmap::FeatureMatches fm(m);
    std::memcpy(fm.data(), matches, m * sizeof(colmap::FeatureMatch));
    colmap::Camera camera1, camera2;   // dummy cameras, as in the reference (:89)
    colmap::TwoViewGeometry::Options opt;
    opt.min_num_inliers = static_cast<size_t>(args_.siftargs.min_num_inliers);
    opt.ransac_options.max_error = args_.siftargs.max_error;
    opt.ransac_options.confidence = args_.siftargs.confidence;
    opt.ransac_options.min_num_trials = static_cast<size_t>(args_.siftargs.min_num_trials);
    opt.ransac_options.max_num_trials = static_cast<size_t>(args_.siftargs.max_num_trials);
    opt.ransac_options.min_inlier_ratio = args_.siftargs.min_inlier_ratio;
    colmap::TwoViewGeometry g;
    const auto p1 = colmap::FeatureKeypointsToPointsVector(c1), p2 = colmap::FeatureKeypointsToPointsVector(c2);
    if (args_.siftargs.multiple_models) g.EstimateMultiple(camera1, p1, camera2, p2, fm, opt);
    else g.Estimate(camera1, p1, camera2, p2, fm, opt);
    out.config = g.config;
    std::memcpy(out.E, g.E.data(), 72); std::memcpy(out.F, g.F.data(), 72); std::memcpy(out.H, g.H.data(), 72);
    std::memcpy(out.qvec, g.qvec.data(), 32); std::memcpy(out.tvec, g.tvec.data(), 24);
    out.tri_angle = g.tri_angle;
    out.inlier_matches.resize(g.inlier_matches.size());
    std::memcpy(out.inlier_matches.data(), g.inlier_matches.data(), 8 * g.inlier_matches.size());
#else
    (void)kp1; (void)kp2;
    out.inlier_matches.assign(reinterpret_cast<const FeatureMatch*>(matches), reinterpret_cast<const FeatureMatch*>(matches) + m);
#endif
    return out;
  }

  smb_proto::SequentialMatchingArgs args_;
  smb_options opts_{};
  smb_tvg_options tvg_opts_{};
  SharedGpu* gpu_ = nullptr;
  bool verbose_ = false;
  bool gpu_verify_ = false;
};

}  // namespace

// Identical registration text to sequential_matching.cc:193-205 (the kernel stays DeviceType::CPU so that
// feature_matching.py, which requests no device, is unchanged; the kernel owns the GPU internally).
REGISTER_OP(SequentialMatchingCPU)
    .stencil()
    .input("image_ids")
    .input("keypoints")
    .input("descriptors")
    .output("pair_image_ids")
    .output("two_view_geometries")
    .protobuf_name("featureMatchingArgs");

REGISTER_KERNEL(SequentialMatchingCPU, SequentialMatchingB200Kernel)
    .device(scanner::DeviceType::CPU)
    .batch()
    .num_devices(1);
