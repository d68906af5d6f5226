#pragma once
#include <functional>
#include <map>
#include <memory>
#include "scanner/util/common.h"
#include "scanner/util/memory.h"

namespace scanner {

struct KernelConfig {
  std::vector<DeviceHandle> devices;
  std::vector<std::string> input_columns;
  std::vector<std::string> output_columns;
  std::vector<u8> args;  // serialized protobuf kernel args (sequential_matching.cc:37-38)
  i32 node_id = 0;
};

class BaseKernel {
 public:
  explicit BaseKernel(const KernelConfig&) {}
  virtual ~BaseKernel() {}
  // Scanner calls reset() when the next rows are not contiguous with the previous ones, and new_stream() when
  // the kernel instance is handed another table / job [ext]: whatever a kernel cached about earlier rows is
  // stale from then on.
  virtual void reset() {}
  virtual void new_stream(const std::vector<u8>& /*args*/) {}
};

// sequential_matching.cc:27-33,103-108: execute(const StenciledBatchedElements&, BatchedElements&)
class StenciledBatchedKernel : public BaseKernel {
 public:
  explicit StenciledBatchedKernel(const KernelConfig& c) : BaseKernel(c) {}
  virtual void execute(const StenciledBatchedElements& input_columns, BatchedElements& output_columns) = 0;
};

class VideoKernel {
 public:
  virtual ~VideoKernel() {}
};

// --- registration (REGISTER_OP / REGISTER_KERNEL builder chains, sequential_matching.cc:193-205)
struct OpRegistration {
  std::string name;
  bool stencil = false;
  std::vector<std::string> inputs, outputs;
  std::string protobuf;
  OpRegistration& stencil_() { stencil = true; return *this; }
};
struct KernelRegistration {
  std::string op;
  DeviceType device = DeviceType::CPU;
  bool batched = false;
  int num_devices = 1;
  std::function<StenciledBatchedKernel*(const KernelConfig&)> factory;
};
inline std::map<std::string, OpRegistration>& op_registry() { static std::map<std::string, OpRegistration> r; return r; }
inline std::map<std::string, KernelRegistration>& kernel_registry() { static std::map<std::string, KernelRegistration> r; return r; }

class OpBuilder {
 public:
  explicit OpBuilder(const std::string& n) { reg_.name = n; }
  OpBuilder& stencil() { reg_.stencil = true; commit(); return *this; }
  OpBuilder& input(const std::string& c) { reg_.inputs.push_back(c); commit(); return *this; }
  OpBuilder& output(const std::string& c) { reg_.outputs.push_back(c); commit(); return *this; }
  OpBuilder& protobuf_name(const std::string& p) { reg_.protobuf = p; commit(); return *this; }
 private:
  void commit() { op_registry()[reg_.name] = reg_; }
  OpRegistration reg_;
};
class KernelBuilder {
 public:
  KernelBuilder(const std::string& op, std::function<StenciledBatchedKernel*(const KernelConfig&)> f) { reg_.op = op; reg_.factory = f; commit(); }
  KernelBuilder& device(DeviceType d) { reg_.device = d; commit(); return *this; }
  KernelBuilder& batch() { reg_.batched = true; commit(); return *this; }
  KernelBuilder& num_devices(int n) { reg_.num_devices = n; commit(); return *this; }
 private:
  void commit() { kernel_registry()[reg_.op] = reg_; }
  KernelRegistration reg_;
};
}  // namespace scanner

#define REGISTER_OP(name) static ::scanner::OpBuilder smb_op_builder_##name __attribute__((unused)) = ::scanner::OpBuilder(#name)
#define REGISTER_KERNEL(name, kernel)                                                                  \
  static ::scanner::KernelBuilder smb_kernel_builder_##name __attribute__((unused)) = ::scanner::KernelBuilder( \
      #name, [](const ::scanner::KernelConfig& c) -> ::scanner::StenciledBatchedKernel* { return new kernel(c); })
