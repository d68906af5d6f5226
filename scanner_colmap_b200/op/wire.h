// Row (de)serialisation of the SequentialMatchingCPU op: the five element formats of
// /root/reference/integration/op_cpp/io.cc that the matcher's op reads or writes, re-implemented as
// bounds-checked, zero-copy views (the reference memcpy-deserialises every stencil entry of every row:
// read_matrix_from_element, io.cc:181-194).  Byte layouts are identical; see tests/test_wire_formats.py.
//
//   image id         8-byte size_t written by prepare_image.cc:17; the op reads the low 4 bytes as
//                    colmap::image_t (io.cc:54-56 via sequential_matching.cc:116-117)
//   keypoints        [size_t n][n x FeatureKeypoint{float x, y, a11, a12, a21, a22}]           io.cc:115-123,151-162
//   descriptors      [size_t rows][size_t cols][rows*cols uint8 row-major]                      io.cc:181-194,198-212
//   pair_image_ids   [size_t n][n x uint32 image_id2]                                            io.cc:151-176
//   two_view_geoms   [size_t total_bytes][int32 n] n x {int32 config; double E[9], F[9], H[9] (col-major);
//                    double qvec[4], tvec[3], tri_angle; size_t m; m x FeatureMatch{u32,u32}}    io.cc:224-297
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace smb_wire {

struct FeatureKeypoint { float x, y, a11, a12, a21, a22; };   // colmap::FeatureKeypoint (24 B)
struct FeatureMatch { uint32_t idx1, idx2; };                 // colmap::FeatureMatch (8 B)
static_assert(sizeof(FeatureKeypoint) == 24 && sizeof(FeatureMatch) == 8, "POD layouts");

// The serialised members of colmap::TwoViewGeometry, in io.cc:283-293 order.
struct TwoViewGeometry {
  int32_t config = 0;                       // TwoViewGeometry::UNDEFINED
  double E[9] = {0}, F[9] = {0}, H[9] = {0};  // Eigen::Matrix3d storage order (column-major)
  double qvec[4] = {0}, tvec[3] = {0};
  double tri_angle = 0;
  std::vector<FeatureMatch> inlier_matches;
};
constexpr size_t kTvgFixedBytes = 4 + 3 * 72 + 32 + 24 + 8;   // 284

struct WireError : std::runtime_error { using std::runtime_error::runtime_error; };

inline uint32_t read_image_id(const uint8_t* buf, size_t size) {
  if (size < 4) throw WireError("image id element shorter than 4 bytes");
  uint32_t id;
  std::memcpy(&id, buf, 4);   // low half of the little-endian size_t
  return id;
}

struct DescriptorView { const uint8_t* data; size_t rows, cols; };
inline DescriptorView view_descriptors(const uint8_t* buf, size_t size) {
  if (size < 16) throw WireError("descriptor element shorter than its header");
  uint64_t rows, cols;
  std::memcpy(&rows, buf, 8);
  std::memcpy(&cols, buf + 8, 8);
  if (rows && cols != 128) throw WireError("descriptor matrix must have 128 columns, got " + std::to_string(cols));
  if (rows > (size - 16) / (cols ? cols : 1)) throw WireError("descriptor element truncated");
  return {buf + 16, (size_t)rows, (size_t)cols};
}

struct KeypointView { const FeatureKeypoint* data; size_t n; };
inline KeypointView view_keypoints(const uint8_t* buf, size_t size) {
  if (size < 8) throw WireError("keypoint element shorter than its header");
  uint64_t n;
  std::memcpy(&n, buf, 8);
  if (n > (size - 8) / sizeof(FeatureKeypoint)) throw WireError("keypoint element truncated");
  return {reinterpret_cast<const FeatureKeypoint*>(buf + 8), (size_t)n};
}

inline size_t pair_ids_bytes(size_t n) { return 8 + 4 * n; }
inline void write_pair_ids(uint8_t* dst, const std::vector<uint32_t>& ids) {
  const uint64_t n = ids.size();
  std::memcpy(dst, &n, 8);
  if (n) std::memcpy(dst + 8, ids.data(), 4 * n);
}

inline size_t tvg_list_bytes(const std::vector<TwoViewGeometry>& l) {
  size_t b = 8 + 4;
  for (const auto& t : l) b += kTvgFixedBytes + 8 + 8 * t.inlier_matches.size();
  return b;
}
inline void write_tvg_list(uint8_t* dst, const std::vector<TwoViewGeometry>& l) {
  uint8_t* p = dst;
  const uint64_t total = tvg_list_bytes(l);
  const int32_t n = (int32_t)l.size();
  std::memcpy(p, &total, 8); p += 8;
  std::memcpy(p, &n, 4); p += 4;
  for (const auto& t : l) {
    std::memcpy(p, &t.config, 4); p += 4;
    std::memcpy(p, t.E, 72); p += 72;
    std::memcpy(p, t.F, 72); p += 72;
    std::memcpy(p, t.H, 72); p += 72;
    std::memcpy(p, t.qvec, 32); p += 32;
    std::memcpy(p, t.tvec, 24); p += 24;
    std::memcpy(p, &t.tri_angle, 8); p += 8;
    const uint64_t m = t.inlier_matches.size();
    std::memcpy(p, &m, 8); p += 8;
    if (m) std::memcpy(p, t.inlier_matches.data(), 8 * m);
    p += 8 * m;
  }
}
inline std::vector<TwoViewGeometry> read_tvg_list(const uint8_t* buf, size_t size) {
  if (size < 12) throw WireError("two_view_geometries element shorter than its header");
  uint64_t total; int32_t n;
  std::memcpy(&total, buf, 8);
  std::memcpy(&n, buf + 8, 4);
  if (total != size || n < 0) throw WireError("two_view_geometries length check failed (io.cc:249 assert)");
  const uint8_t* p = buf + 12; const uint8_t* end = buf + size;
  std::vector<TwoViewGeometry> l((size_t)n);
  for (auto& t : l) {
    if ((size_t)(end - p) < kTvgFixedBytes + 8) throw WireError("two_view_geometries truncated");
    std::memcpy(&t.config, p, 4); p += 4;
    std::memcpy(t.E, p, 72); p += 72;
    std::memcpy(t.F, p, 72); p += 72;
    std::memcpy(t.H, p, 72); p += 72;
    std::memcpy(t.qvec, p, 32); p += 32;
    std::memcpy(t.tvec, p, 24); p += 24;
    std::memcpy(&t.tri_angle, p, 8); p += 8;
    uint64_t m; std::memcpy(&m, p, 8); p += 8;
    if (m > (size_t)(end - p) / 8) throw WireError("two_view_geometries match list truncated");
    t.inlier_matches.resize((size_t)m);
    if (m) std::memcpy(t.inlier_matches.data(), p, 8 * m);
    p += 8 * m;
  }
  if (p != end) throw WireError("two_view_geometries trailing bytes");
  return l;
}

}  // namespace smb_wire
