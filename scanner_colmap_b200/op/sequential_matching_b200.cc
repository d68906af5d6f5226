// Drop-in replacement for the Scanner op of
//   /root/reference/integration/op_cpp/sequential_matching.cc
// Same registered op name (SequentialMatchingCPU), same input / output columns, same kernel-args protobuf,
// same serialized rows (io.cc) -- so integration/feature_matching.py and the downstream incremental_mapping
// stage (incremental_mapping.cc:245-260) consume it unchanged.  What changes: the call
//   colmap::MatchSiftFeaturesCPU(sift_options_, descriptors1, descriptors2, &featureMatches)   (:154)
// is served by the B200 library behind include/smb.h, for ALL pairs of ALL rows of the batch in one call,
// with descriptors uploaded once per image (the reference re-deserialises every stencil entry per row and
// only looks at batch item 0, :106-108).
//
// Two-view geometry verification (:84-101, :157-178) stays on the CPU and stays COLMAP's: with
// -DSMB_WITH_COLMAP the unchanged colmap::TwoViewGeometry::Estimate[Multiple] runs on the GPU matches; without
// COLMAP (this image) the verifier is a pass-through (config = UNDEFINED, inlier_matches = raw matches) so the
// row format, the pair enumeration and the min_num_inliers filter can still be exercised end to end.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <unordered_set>

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

class SequentialMatchingB200Kernel : public scanner::StenciledBatchedKernel, public scanner::VideoKernel {
 public:
  explicit SequentialMatchingB200Kernel(const scanner::KernelConfig& config) : scanner::StenciledBatchedKernel(config) {
    // sequential_matching.cc:36-76: empty args == proto2 defaults (what feature_matching.py sends)
    smb_proto::parse(config.args.data(), config.args.size(), args_);
    smb_options o;
    smb_default_options(&o);
    o.max_ratio = args_.siftargs.max_ratio;
    o.max_distance = args_.siftargs.max_distance;
    o.cross_check = args_.siftargs.cross_check ? 1 : 0;
    o.max_num_matches = args_.siftargs.max_num_matches;
    int device = 0;
    if (const char* e = std::getenv("SMB_DEVICE")) device = std::atoi(e);
    else if (args_.siftargs.gpu_index != "-1") device = std::atoi(args_.siftargs.gpu_index.c_str());
    verbose_ = std::getenv("SMB_OP_VERBOSE") != nullptr;  // the reference printf's per pair (:130-134,150,171)
    const int rc = smb_create(device, &o, &h_);
    SMB_CHECK(rc == SMB_OK, "smb_create: %s", smb_last_error(nullptr));
  }
  ~SequentialMatchingB200Kernel() override { smb_destroy(h_); }

  void execute(const scanner::StenciledBatchedElements& input_cols, scanner::BatchedElements& output_cols) override {
    SMB_CHECK(input_cols.size() >= 3 && output_cols.size() >= 2, "expected 3 input and 2 output columns");
    const size_t batch = input_cols[0].size();
    // ---- 1. decode every row's stencil; collect the distinct images of the batch
    struct Row { std::vector<uint32_t> ids; std::vector<uint32_t> partners; std::vector<size_t> partner_stencil; };
    std::vector<Row> rows(batch);
    std::vector<uint32_t> new_ids;
    std::vector<const uint8_t*> new_desc;
    std::vector<size_t> new_n;
    std::unordered_set<uint32_t> used;
    std::vector<uint32_t> pairs;
    for (size_t b = 0; b < batch; ++b) {
      const scanner::Elements& id_st = input_cols[0][b];
      const scanner::Elements& desc_st = input_cols[2][b];
      SMB_CHECK(id_st.size() == desc_st.size() && id_st.size() == input_cols[1][b].size(), "stencil sizes differ");
      Row& r = rows[b];
      for (size_t s = 0; s < id_st.size(); ++s) {
        const uint32_t id = smb_wire::read_image_id(id_st[s].buffer, id_st[s].size);
        r.ids.push_back(id);
        if (used.insert(id).second && !smb_has_image(h_, id)) {
          const smb_wire::DescriptorView d = smb_wire::view_descriptors(desc_st[s].buffer, desc_st[s].size);
          new_ids.push_back(id);
          new_desc.push_back(d.data);
          new_n.push_back(d.rows);
        }
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
    // ---- 2. descriptor cache: drop images no row of this batch names, upload the new ones (one wait)
    for (auto it = cached_.begin(); it != cached_.end();) {
      if (!used.count(*it)) {
        SMB_CHECK(smb_evict_image(h_, *it) == SMB_OK, "%s", smb_last_error(h_));
        it = cached_.erase(it);
      } else ++it;
    }
    if (!new_ids.empty()) {
      const int rc = smb_put_images(h_, new_ids.data(), new_desc.data(), new_n.data(), new_ids.size(), SMB_DESC_DIM);
      SMB_CHECK(rc == SMB_OK, "smb_put_images: %s", smb_last_error(h_));
      cached_.insert(new_ids.begin(), new_ids.end());
    }
    // ---- 3. every pair of the batch in one GPU call
    smb_result* res = nullptr;
    const int rc = smb_match_pairs(h_, pairs.data(), pairs.size() / 2, &res);
    SMB_CHECK(rc == SMB_OK, "smb_match_pairs: %s", smb_last_error(h_));
    // ---- 4. verify + serialise, one output row per batch item (the reference emits one, for item 0 only)
    size_t p = 0;
    for (size_t b = 0; b < batch; ++b) {
      const Row& r = rows[b];
      std::vector<TwoViewGeometry> tvgs;
      tvgs.reserve(r.partners.size());
      for (size_t k = 0; k < r.partners.size(); ++k, ++p) {
        size_t m = 0;
        const smb_match* matches = smb_result_matches(res, p, &m);
        TwoViewGeometry tvg = verify(matches, m, input_cols[1][b][0], input_cols[1][b][r.partner_stencil[k]]);
        if (verbose_) std::printf("View geometry for #%u and #%u has %zu inliers\n", r.ids[0], r.partners[k], tvg.inlier_matches.size());
        // sequential_matching.cc:173-178: too few inliers -> default-constructed TwoViewGeometry
        if (tvg.inlier_matches.size() < static_cast<size_t>(args_.siftargs.min_num_inliers)) tvg = TwoViewGeometry();
        tvgs.push_back(std::move(tvg));
      }
      const size_t n0 = smb_wire::pair_ids_bytes(r.partners.size());
      scanner::u8* b0 = scanner::new_buffer(scanner::CPU_DEVICE, n0);
      smb_wire::write_pair_ids(b0, r.partners);
      scanner::insert_element(output_cols[0], b0, n0);                  // io.cc:151-176
      const size_t n1 = smb_wire::tvg_list_bytes(tvgs);
      scanner::u8* b1 = scanner::new_buffer(scanner::CPU_DEVICE, n1);
      smb_wire::write_tvg_list(b1, tvgs);
      scanner::insert_element(output_cols[1], b1, n1);                  // io.cc:256-304
    }
    smb_result_release(h_, res);
  }

 private:
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
    out.inlier_matches.resize(m);
    if (m) std::memcpy(out.inlier_matches.data(), matches, m * sizeof(FeatureMatch));
#endif
    return out;
  }

  smb_proto::SequentialMatchingArgs args_;
  smb_handle* h_ = nullptr;
  std::set<uint32_t> cached_;
  bool verbose_ = false;
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
