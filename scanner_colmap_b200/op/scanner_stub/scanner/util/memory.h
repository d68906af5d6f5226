#pragma once
#include "scanner/util/common.h"

namespace scanner {
// Scanner owns output buffers: kernels allocate with new_buffer and hand them over with insert_element
// (io.cc:157-161, 173-176, 272, 302-303).  The stub uses malloc; the harness frees.
inline u8* new_buffer(const DeviceHandle&, size_t size) { return static_cast<u8*>(std::malloc(size ? size : 1)); }
inline void delete_buffer(const DeviceHandle&, u8* p) { std::free(p); }
inline void insert_element(Elements& col, u8* buffer, size_t size) { col.emplace_back(buffer, size); }
inline void memcpy_buffer(u8* dst, const DeviceHandle&, const u8* src, const DeviceHandle&, size_t n) { std::memcpy(dst, src, n); }
}  // namespace scanner
