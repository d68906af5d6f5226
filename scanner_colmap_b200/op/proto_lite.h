// Decoder for the two protobuf messages the op's kernel args use (colmap.proto, proto2):
//   SequentialMatchingArgs { bool loop_detection=1; int32 overlap=2; bool quadratic_overlap=3; siftFeatureMatchingArgs siftArgs=4; }
//   siftFeatureMatchingArgs { 14 fields, defaults colmap.proto:7-48 }
// Neither protoc nor libprotobuf is available in this image; with them present the generated colmap.pb.h can
// be used instead (-DSMB_WITH_PROTOBUF).  Unknown fields are skipped, absent fields keep their proto2 defaults,
// malformed input makes parse() return false (the reference ignores ParseFromArray's result; empty args ==
// all defaults, which is what feature_matching.py sends).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace smb_proto {

struct SiftFeatureMatchingArgs {
  bool use_gpu = false;              // 1
  std::string gpu_index = "-1";      // 2
  double max_ratio = 0.8;            // 3
  double max_distance = 0.7;         // 4
  bool cross_check = true;           // 5
  int32_t max_num_matches = 32768;   // 6
  float max_error = 4.0f;            // 7
  double confidence = 0.999;         // 8
  int32_t min_num_trials = 30;       // 9
  int32_t max_num_trials = 10000;    // 10
  double min_inlier_ratio = 0.25;    // 11
  int32_t min_num_inliers = 15;      // 12
  bool multiple_models = false;      // 13
  bool guided_matching = false;      // 14
};
struct SequentialMatchingArgs {
  bool loop_detection = false;       // 1
  int32_t overlap = 10;              // 2
  bool quadratic_overlap = false;    // 3
  SiftFeatureMatchingArgs siftargs;  // 4
};

namespace detail {
struct Reader {
  const uint8_t* p; const uint8_t* end; bool ok = true;
  bool varint(uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 64 && p < end; shift += 7) {
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return true;
    }
    return ok = false;
  }
  bool fixed(void* dst, size_t n) {
    if ((size_t)(end - p) < n) return ok = false;
    std::memcpy(dst, p, n); p += n; return true;
  }
  bool skip(uint32_t wire) {
    uint64_t v;
    switch (wire) {
      case 0: return varint(v);
      case 1: { uint64_t d; return fixed(&d, 8); }
      case 2: if (!varint(v) || (uint64_t)(end - p) < v) return ok = false; p += v; return true;
      case 5: { uint32_t d; return fixed(&d, 4); }
      default: return ok = false;
    }
  }
};
inline bool parse_sift(const uint8_t* data, size_t n, SiftFeatureMatchingArgs& a) {
  Reader r{data, data + n};
  while (r.p < r.end && r.ok) {
    uint64_t key, v; if (!r.varint(key)) break;
    const uint32_t field = (uint32_t)(key >> 3), wire = (uint32_t)(key & 7);
    auto dbl = [&](double& d) { if (wire == 1) r.fixed(&d, 8); else r.skip(wire); };
    auto i32 = [&](int32_t& d) { if (wire == 0) { if (r.varint(v)) d = (int32_t)v; } else r.skip(wire); };
    auto bl = [&](bool& d) { if (wire == 0) { if (r.varint(v)) d = v != 0; } else r.skip(wire); };
    switch (field) {
      case 1: bl(a.use_gpu); break;
      case 2: if (wire == 2 && r.varint(v) && (uint64_t)(r.end - r.p) >= v) { a.gpu_index.assign((const char*)r.p, (size_t)v); r.p += v; } else r.ok = false; break;
      case 3: dbl(a.max_ratio); break;
      case 4: dbl(a.max_distance); break;
      case 5: bl(a.cross_check); break;
      case 6: i32(a.max_num_matches); break;
      case 7: if (wire == 5) r.fixed(&a.max_error, 4); else r.skip(wire); break;
      case 8: dbl(a.confidence); break;
      case 9: i32(a.min_num_trials); break;
      case 10: i32(a.max_num_trials); break;
      case 11: dbl(a.min_inlier_ratio); break;
      case 12: i32(a.min_num_inliers); break;
      case 13: bl(a.multiple_models); break;
      case 14: bl(a.guided_matching); break;
      default: r.skip(wire);
    }
  }
  return r.ok;
}
}  // namespace detail

inline bool parse(const uint8_t* data, size_t n, SequentialMatchingArgs& a) {
  detail::Reader r{data, data + n};
  while (r.p < r.end && r.ok) {
    uint64_t key, v; if (!r.varint(key)) break;
    const uint32_t field = (uint32_t)(key >> 3), wire = (uint32_t)(key & 7);
    if (field == 1 && wire == 0) { if (r.varint(v)) a.loop_detection = v != 0; }
    else if (field == 2 && wire == 0) { if (r.varint(v)) a.overlap = (int32_t)v; }
    else if (field == 3 && wire == 0) { if (r.varint(v)) a.quadratic_overlap = v != 0; }
    else if (field == 4 && wire == 2) {
      if (!r.varint(v) || (uint64_t)(r.end - r.p) < v) { r.ok = false; break; }
      if (!detail::parse_sift(r.p, (size_t)v, a.siftargs)) r.ok = false;
      r.p += v;
    } else r.skip(wire);
  }
  return r.ok;
}

}  // namespace smb_proto
