/*
 * smb.h -- C ABI of the B200 (sm_100a) SIFT descriptor matcher ("sift match, B200").
 *
 * Drop-in boundary for ONE call of the reference (garyjyzhang/scanner-colmap):
 *
 *     colmap::MatchSiftFeaturesCPU(sift_options_, descriptors1, descriptors2, &featureMatches);
 *         -- integration/op_cpp/sequential_matching.cc:154
 *
 * as it is driven by the per-row pair loop of the SequentialMatchingCPU Scanner
 * op (sequential_matching.cc:139-181).  The op shim (scanner_colmap_b200/op/)
 * and the Python mirror (scanner_colmap_b200/matcher.py, ctypes) are the only
 * callers.  Plain pointers and sizes only; no exceptions cross this boundary;
 * no global state besides the create-time error string; one CUDA stream set per
 * handle; a handle may be used by one thread at a time (Scanner calls execute()
 * serially per kernel instance), different handles are independent.
 *
 * All functions returning int return SMB_OK (0) or a negative SMB_E* code;
 * smb_last_error() gives the text.  The op shim CHECK_EQ's the code against 0,
 * which matches the reference's abort-on-error convention (io.cc:392-404).
 */
#ifndef SMB_H_
#define SMB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMB_ABI_VERSION 2
#define SMB_DESC_DIM 128 /* SIFT descriptor bytes; FeatureDescriptors cols, io.cc:181-194 */

enum {
  SMB_OK = 0,
  SMB_EINVAL = -1,    /* bad argument (null pointer, d != 128, unknown image id ...) */
  SMB_ECUDA = -2,     /* CUDA runtime / driver error; text in smb_last_error */
  SMB_ENOMEM = -3,    /* host or device allocation failed */
  SMB_ENODEVICE = -4, /* no sm_100 device / wrong architecture: there is NO CPU fallback */
  SMB_ECAPACITY = -5  /* caller buffer too small */
};

/* Which device kernel computes the dot-product tiles.  The product library (libsmb.so) contains exactly one
 * engine; SMB_ENGINE_DP4A exists only in the test build (libsmb_test.so, compiled with -DSMB_TEST_ENGINES) where
 * it cross-checks the tensor-core path on CUDA cores.  libsmb.so rejects it with SMB_EINVAL. */
enum {
  SMB_ENGINE_TCGEN05 = 0, /* TMA + tcgen05.mma kind::i8 + TMEM */
  SMB_ENGINE_DP4A = 1     /* test builds only */
};

/* Mirrors the fields of colmap::SiftMatchingOptions that MatchSiftFeaturesCPU reads,
 * filled by the op from siftFeatureMatchingArgs (sequential_matching.cc:40-55,
 * defaults colmap.proto:14-24).  max_ratio / max_distance are the double options
 * narrowed to float exactly where COLMAP narrows them (the FindBestMatches call). */
typedef struct smb_options {
  double max_ratio;       /* colmap.proto:14  default 0.8 */
  double max_distance;    /* colmap.proto:17  default 0.7 */
  int32_t cross_check;    /* colmap.proto:20  default true */
  int32_t max_num_matches; /* colmap.proto:23 default 32768; accepted and, like the reference's CPU
                              path, NOT applied (only COLMAP's GPU matcher truncates) */
  int32_t engine;         /* SMB_ENGINE_* */
  int32_t profile;        /* non-zero: record CUDA events per kernel, see smb_get_timing */
} smb_options;

/* == colmap::FeatureMatch {point2D_idx1, point2D_idx2}, the 8-byte POD that io.cc:267-289
 * memcpy-serialises into the two_view_geometries row. */
typedef struct smb_match {
  uint32_t idx1;
  uint32_t idx2;
} smb_match;

typedef struct smb_handle smb_handle;
typedef struct smb_result smb_result;

/* Device timings of the most recent completed match call (milliseconds, CUDA events on the handle's stream; the
 * *_ms fields are valid only when options.profile != 0). */
typedef struct smb_timing {
  float total_ms;       /* first kernel of the call to the last one (results are in host memory when it ends) */
  float score_ms;       /* score_tcgen05_kernel only, summed over the call's sub-batches */
  float runner_up_ms;   /* runner_up_kernel, summed (kernel time: with several sub-batches these two run on a second */
  float decide_ms;      /* stream underneath the next sub-batch's score kernel); decide_kernel writes the matches
                           straight into pinned host memory */
  uint32_t score_launches;
  uint32_t total_launches; /* every kernel of this library launched by the call */
  uint32_t sub_batches;
  uint32_t plan_uploaded;  /* 0: the device-side plan of the previous call was reused (same pairs, same pool rows) */
  uint64_t candidates;  /* score-matrix entries that survived the integer pre-filter */
  uint64_t ops;         /* 2 * sum(n1 * n2) * 128 over the call's pairs */
  float cta_busy_max_over_mean; /* load balance of the persistent score CTAs: sum over the call's score launches of the
                                   busiest CTA's time / sum of the mean CTA time (1.0 = perfectly balanced) */
  float pad_;
} smb_timing;

void smb_default_options(smb_options* opts);
int smb_abi_version(void);

/* Create a matcher bound to one CUDA device.  Fails with SMB_ENODEVICE when the device is not
 * compute capability 10.x: there is no fallback path. */
int smb_create(int cuda_device, const smb_options* opts, smb_handle** out);
void smb_destroy(smb_handle* h);
/* h may be NULL to read the error of a failed smb_create (thread-local). */
const char* smb_last_error(const smb_handle* h);

/* Replace the options (thresholds are re-derived; cached images stay). */
int smb_set_options(smb_handle* h, const smb_options* opts);

/* Descriptor cache.  Replaces the per-row re-deserialisation of
 * read_matrix_from_element<FeatureDescriptors> (io.cc:181-194, call sites
 * sequential_matching.cc:120-121): an image is uploaded once per handle and reused by
 * every pair that names it.  desc = n x 128 uint8 row-major host memory (pageable or
 * pinned); the call returns after the bytes have been consumed.  Re-putting an existing
 * id replaces it. */
int smb_put_image(smb_handle* h, uint32_t image_id, const uint8_t* desc, size_t n, size_t d);
/* Batched form: all copies are queued, one wait at the end (the op uses it per stencil window). */
int smb_put_images(smb_handle* h, const uint32_t* image_ids, const uint8_t* const* descs, const size_t* ns,
                   size_t count, size_t d);
/* Asynchronous form: the copies are queued on a dedicated upload stream and the call returns at once; the
 * caller keeps the (pinned) buffers unchanged until smb_synchronize() or until a match call that names the
 * images has returned.  A match call takes its pairs in the order their images land and its score kernel waits,
 * item by item on the device, for the upload an item depends on: ONE launch covers a call whose images are still
 * crossing PCIe, and uploading the next images overlaps matching the previous ones. */
int smb_put_images_async(smb_handle* h, const uint32_t* image_ids, const uint8_t* const* descs, const size_t* ns,
                         size_t count, size_t d);
/* Same, from device memory of this or a peer device (halo exchange over NVLink). */
int smb_put_image_device(smb_handle* h, uint32_t image_id, const void* dev_desc, size_t n, size_t d);
int smb_put_images_device(smb_handle* h, const uint32_t* image_ids, const void* const* dev_descs, const size_t* ns,
                          size_t count, size_t d);
/* Asynchronous form of smb_put_images_device: the device buffers are (being) produced by work already queued on
 * `producer_stream` (a cudaStream_t, e.g. the stream an NCCL recv of the halo was launched on; NULL = the legacy
 * default stream).  The library orders its device-to-device copies after that work on its upload stream and
 * returns at once; like smb_put_images_async, a match call waits on the device only for the uploads its own pairs
 * need, so the halo exchange overlaps the matching of every pair that does not touch a halo image.  The buffers
 * must stay valid and unchanged until smb_synchronize() or until a match call naming the images has returned. */
int smb_put_images_device_async(smb_handle* h, const uint32_t* image_ids, const void* const* dev_descs,
                                const size_t* ns, size_t count, size_t d, void* producer_stream);
int smb_has_image(const smb_handle* h, uint32_t image_id);
int smb_evict_image(smb_handle* h, uint32_t image_id);
int smb_clear_images(smb_handle* h);
/* Device address and row count of a cached image (read-only; valid until eviction/put/destroy). */
int smb_image_device_ptr(const smb_handle* h, uint32_t image_id, const void** dev_ptr, size_t* n);

/* Match npairs (image_id1, image_id2) pairs of cached images: the flattened pair loop of
 * sequential_matching.cc:139-181.  Result i holds exactly what MatchSiftFeaturesCPU would have
 * written to featureMatches for pair i: (idx1, idx2) in ascending idx1. */
int smb_match_pairs(smb_handle* h, const uint32_t* pairs /* [npairs][2] */, size_t npairs, smb_result** out);
/* The same call in two halves, so the caller (the op: verification and serialisation of the previous packet) can
 * work while the GPU matches: _begin queues every kernel and returns; smb_result_wait blocks until the matches
 * are in host memory.  At most one call may be in flight per handle; descriptor uploads may be queued meanwhile
 * (smb_put_images_async), evictions take effect for calls begun afterwards. */
int smb_match_pairs_begin(smb_handle* h, const uint32_t* pairs /* [npairs][2] */, size_t npairs, smb_result** out);
int smb_result_wait(smb_handle* h, smb_result* r);

size_t smb_result_num_pairs(const smb_result* r);
/* Matches of pair i; pointer is into pinned host memory owned by the result (the device wrote it directly). */
const smb_match* smb_result_matches(const smb_result* r, size_t i, size_t* count);
size_t smb_result_total_matches(const smb_result* r);
void smb_result_release(smb_handle* h, smb_result* r);

/* One-shot form with the exact shape of the replaced call: host descriptors in, matches out.
 * out must hold min(n1, n2) entries when cross_check, else n1.  Nothing is cached. */
int smb_match_descriptors(smb_handle* h, const uint8_t* desc1, size_t n1, const uint8_t* desc2, size_t n2,
                          smb_match* out, size_t capacity, size_t* count);

int smb_get_timing(const smb_handle* h, smb_timing* t);

/* ---- Two-view geometry verification on the GPU (the step after the matcher in the reference op:
 * verifyTwoViewGeometry -> colmap::TwoViewGeometry::Estimate, sequential_matching.cc:84-101,157-178, which with the
 * reference's dummy cameras is the uncalibrated F / H LORANSAC path).  Statistical contract, not bit parity: COLMAP
 * samples from a thread-local PRNG (DESIGN.md "Two-view geometry"). */
/* smb_tvg_options.flags.  COLMAP's TwoViewGeometry::Options has detect_watermark = true (a geometry whose inliers are
 * one pure image translation is reported as WATERMARK; with the reference's default-constructed cameras,
 * sequential_matching.cc:89, the border-region condition holds for every inlier) -- so detection is ON unless the
 * first bit is set.  The second bit selects TwoViewGeometry::EstimateMultiple (siftMatchingArgs.multiple_models,
 * colmap.proto:45, sequential_matching.cc:94-96): estimate, remove the inliers, repeat until DEGENERATE; watermark
 * models are skipped (multiple_ignore_watermark = true); several models -> config MULTIPLE with all inlier matches. */
#define SMB_TVG_NO_WATERMARK 1
#define SMB_TVG_MULTIPLE_MODELS 2
typedef struct smb_tvg_options {
  int32_t min_num_inliers; /* colmap.proto:41 default 15 */
  int32_t min_num_trials;  /* colmap.proto:32 default 30 */
  int32_t max_num_trials;  /* colmap.proto:33 default 10000 */
  int32_t flags;           /* SMB_TVG_* bits, default 0 */
  double max_error;        /* colmap.proto:26 default 4.0 (pixels) */
  double confidence;       /* colmap.proto:29 default 0.999 */
  double min_inlier_ratio; /* colmap.proto:37 default 0.25 */
  double max_h_inlier_ratio; /* COLMAP's TwoViewGeometry::Options default 0.8 (not in the reference's proto) */
  uint64_t seed;
} smb_tvg_options;
void smb_default_tvg_options(smb_tvg_options* o);

typedef struct smb_tvg {
  int32_t config;           /* colmap::TwoViewGeometry::ConfigurationType: 1 DEGENERATE, 3 UNCALIBRATED, 6 PLANAR_OR_PANORAMIC,
                             * 7 WATERMARK, 8 MULTIPLE (F and H are zero then, as COLMAP leaves them) */
  int32_t num_inliers_f, num_inliers_h;
  int32_t trials_f, trials_h;
  uint32_t inlier_start, inlier_count; /* internal offsets; use smb_result_inliers */
  uint32_t pad_;
  double F[9], H[9];        /* row-major */
} smb_tvg;

/* Keypoint positions (x, y) of a cached image: `n` must equal its descriptor count; `xy` points at the first x,
 * consecutive keypoints are `stride_bytes` apart (24 for colmap::FeatureKeypoint rows, io.cc:115-123; 8 for packed). */
int smb_put_keypoints(smb_handle* h, uint32_t image_id, const float* xy, size_t n, size_t stride_bytes);
/* Verify every pair of `r`, which must be the most recent completed match call of `h`, and whose images must all
 * have keypoints.  Fills per pair a smb_tvg and the inlier matches (ascending idx1, pinned host memory). */
int smb_result_verify(smb_handle* h, smb_result* r, const smb_tvg_options* opts);
int smb_result_tvg(const smb_result* r, size_t i, smb_tvg* out);
const smb_match* smb_result_inliers(const smb_result* r, size_t i, size_t* count);

/* Page-locked host memory for callers that do not link the CUDA runtime themselves (the op stages Scanner's
 * pageable descriptor rows through such a buffer so that smb_put_images_async really is asynchronous). */
int smb_alloc_pinned(size_t bytes, void** out);
void smb_free_pinned(void* p);

/* Integer pre-filter derived from the options and the host libm acosf table: score entries
 * below *min_score cannot change any accept/reject decision ("Filter"). */
int smb_get_filter(const smb_handle* h, int32_t* min_score, int32_t* min_best);

/* CUDA stream the handle launches on (cudaStream_t as void*), for external event timing. */
void* smb_stream(const smb_handle* h);
/* Make `stream` (a cudaStream_t of the same device) wait, on the device, for every upload queued so far by the
 * smb_put_images*_async calls: lets a peer copy / NCCL send read freshly uploaded pool rows (the halo this GPU
 * provides to its neighbour) without any host synchronisation. */
int smb_stream_wait_uploads(smb_handle* h, void* stream);
/* The converse: every kernel of the match calls that follow waits, on the device, for the work queued so far on
 * `stream`.  A rank that only SENDS halo rows must call this after queueing the send: the score kernel is persistent
 * and fills every SM, so a send kernel that has not started yet would otherwise run only after it -- and stall the
 * receiving rank for a whole step.  (Receivers get the same ordering from smb_put_images_device_async.) */
int smb_wait_stream(smb_handle* h, void* stream);
int smb_synchronize(smb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* SMB_H_ */
