// Minimal stand-in for the parts of Scanner's public op API that the SequentialMatching op touches
// (scanner/util/common.h, scanner/util/memory.h, scanner/api/kernel.h, scanner/api/op.h).  Scanner is
// not installable in this image, so the op is compile-checked and driven (tests/, fake dispatch harness)
// against these declarations; with a real Scanner checkout the same op source builds against the real
// headers (-DSMB_WITH_SCANNER).  Written from the call sites in
// /root/reference/integration/op_cpp/{sequential_matching.cc,io.cc}; no Scanner source was available.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace scanner {

using u8 = uint8_t;
using i32 = int32_t;
using i64 = int64_t;

enum class DeviceType { CPU = 0, GPU = 1 };

struct DeviceHandle {
  DeviceType type;
  int id;
  bool operator==(const DeviceHandle& o) const { return type == o.type && id == o.id; }
};
static const DeviceHandle CPU_DEVICE = {DeviceType::CPU, 0};

// One table cell handed to / produced by a kernel (io.cc:67-69 reads element.buffer; io.cc:161 fills it).
struct Element {
  u8* buffer = nullptr;
  size_t size = 0;
  bool is_frame = false;
  Element() = default;
  Element(u8* b, size_t s) : buffer(b), size(s) {}
};
using Elements = std::vector<Element>;
using BatchedElements = std::vector<Elements>;                        // [column][batch item]
using StenciledElements = std::vector<Elements>;                      // [column][stencil]
using StenciledBatchedElements = std::vector<std::vector<Elements>>;  // [column][batch item][stencil]

}  // namespace scanner
