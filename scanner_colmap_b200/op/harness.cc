// C entry points that let Python (ctypes) stand in for Scanner's evaluate worker: look the op up in the
// registry, build a kernel from serialized args, feed it stencilled batches of serialized elements, read the
// output elements back.  Test infrastructure for the op; the product is libsequential_matching.so itself.
#include <cstring>
#include <memory>
#include "scanner/api/kernel.h"
#include "proto_lite.h"
#include "wire.h"

extern "C" {

struct smb_op_kernel { std::unique_ptr<scanner::StenciledBatchedKernel> k; scanner::BatchedElements out; };

int smb_op_registered(const char* op, char* desc, size_t cap) {
  auto it = scanner::op_registry().find(op);
  auto kt = scanner::kernel_registry().find(op);
  if (it == scanner::op_registry().end() || kt == scanner::kernel_registry().end()) return 0;
  std::string s = it->second.name + "|stencil=" + (it->second.stencil ? "1" : "0") + "|in=";
  for (auto& c : it->second.inputs) s += c + ",";
  s += "|out=";
  for (auto& c : it->second.outputs) s += c + ",";
  s += "|proto=" + it->second.protobuf + "|device=" + (kt->second.device == scanner::DeviceType::CPU ? "CPU" : "GPU") +
       "|batch=" + (kt->second.batched ? "1" : "0") + "|num_devices=" + std::to_string(kt->second.num_devices);
  if (desc && cap) { std::strncpy(desc, s.c_str(), cap - 1); desc[cap - 1] = 0; }
  return 1;
}

smb_op_kernel* smb_op_new_kernel(const char* op, const uint8_t* args, size_t nargs) {
  auto kt = scanner::kernel_registry().find(op);
  if (kt == scanner::kernel_registry().end()) return nullptr;
  scanner::KernelConfig cfg;
  cfg.args.assign(args, args + nargs);
  auto* h = new smb_op_kernel();
  h->k.reset(kt->second.factory(cfg));
  return h;
}
void smb_op_delete_kernel(smb_op_kernel* h) {
  if (!h) return;
  for (auto& col : h->out) for (auto& e : col) scanner::delete_buffer(scanner::CPU_DEVICE, e.buffer);
  delete h;
}

void smb_op_reset(smb_op_kernel* h) { h->k->reset(); }
void smb_op_new_stream(smb_op_kernel* h) { h->k->new_stream(std::vector<scanner::u8>()); }

// inputs: ncols x batch x stencil element pointers/sizes flattened in that order
int smb_op_execute(smb_op_kernel* h, size_t ncols, size_t batch, size_t stencil, const uint8_t* const* bufs, const size_t* sizes,
                   size_t nout) {
  scanner::StenciledBatchedElements in(ncols, std::vector<scanner::Elements>(batch, scanner::Elements(stencil)));
  size_t x = 0;
  for (size_t c = 0; c < ncols; ++c)
    for (size_t b = 0; b < batch; ++b)
      for (size_t s = 0; s < stencil; ++s, ++x) in[c][b][s] = scanner::Element(const_cast<uint8_t*>(bufs[x]), sizes[x]);
  for (auto& col : h->out) for (auto& e : col) scanner::delete_buffer(scanner::CPU_DEVICE, e.buffer);
  h->out.assign(nout, scanner::Elements());
  h->k->execute(in, h->out);
  return 0;
}
size_t smb_op_output_count(const smb_op_kernel* h, size_t col) { return h->out[col].size(); }
const uint8_t* smb_op_output(const smb_op_kernel* h, size_t col, size_t i, size_t* size) {
  *size = h->out[col][i].size;
  return h->out[col][i].buffer;
}

// ---- wire-format probes (no GPU needed): the C++ writers/readers against Python-built bytes
size_t smb_wire_tvg_roundtrip(const uint8_t* buf, size_t size, uint8_t* out, size_t cap) {
  try {
    auto l = smb_wire::read_tvg_list(buf, size);
    const size_t n = smb_wire::tvg_list_bytes(l);
    if (n > cap) return 0;
    smb_wire::write_tvg_list(out, l);
    return n;
  } catch (const smb_wire::WireError&) { return 0; }
}
size_t smb_wire_pair_ids(const uint32_t* ids, size_t n, uint8_t* out, size_t cap) {
  std::vector<uint32_t> v(ids, ids + n);
  if (smb_wire::pair_ids_bytes(n) > cap) return 0;
  smb_wire::write_pair_ids(out, v);
  return smb_wire::pair_ids_bytes(n);
}
int smb_wire_descriptor_view(const uint8_t* buf, size_t size, size_t* rows, size_t* cols, size_t* offset) {
  try {
    auto d = smb_wire::view_descriptors(buf, size);
    *rows = d.rows; *cols = d.cols; *offset = (size_t)(d.data - buf);
    return 1;
  } catch (const smb_wire::WireError&) { return 0; }
}
uint32_t smb_wire_image_id(const uint8_t* buf, size_t size) { return smb_wire::read_image_id(buf, size); }

int smb_proto_parse(const uint8_t* buf, size_t n, double* max_ratio, double* max_distance, int* cross_check, int* max_num_matches,
                    int* min_num_inliers, int* overlap, float* max_error) {
  smb_proto::SequentialMatchingArgs a;
  const bool ok = smb_proto::parse(buf, n, a);
  *max_ratio = a.siftargs.max_ratio; *max_distance = a.siftargs.max_distance; *cross_check = a.siftargs.cross_check;
  *max_num_matches = a.siftargs.max_num_matches; *min_num_inliers = a.siftargs.min_num_inliers; *overlap = a.overlap;
  *max_error = a.siftargs.max_error;
  return ok ? 1 : 0;
}

}  // extern "C"
