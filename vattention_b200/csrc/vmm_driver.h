// VMM driver layer: the handful of CUDA *driver* entry points the allocator
// needs (reference: vattention/cudaInternal.h:15-94, vtensor.h:37), behind an
// interface with two implementations:
//   CudaVmmDriver  -- libcuda.so.1 resolved with dlopen/dlsym at first use, so
//                     libvattn_b200.so itself loads on a machine without a
//                     driver (the CPU test box).
//   MockVmmDriver  -- records calls and hands out fake VAs / handle ids; lets
//                     the bookkeeping be checked bit-exactly on CPU.
#pragma once
#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace vattn {

using u64 = uint64_t;

enum DriverOp : u64 {
  OP_RESERVE = 1,
  OP_CREATE = 2,
  OP_MAP = 3,
  OP_SET_ACCESS = 4,
  OP_UNMAP = 5,
  OP_RELEASE = 6,
  OP_ADDR_FREE = 7,
};

struct DriverLogRecord {
  u64 op, va, size, handle;
};

class VmmDriver {
 public:
  virtual ~VmmDriver() = default;
  // cuInit + current-context check + granularity query (cudaInternal.h:15-35).
  // Returns the minimum physical allocation granularity for `device`.
  virtual u64 init(int device) = 0;
  virtual u64 reserve(u64 size, u64 alignment) = 0;            // cuMemAddressReserve
  virtual u64 create(u64 size) = 0;                             // cuMemCreate (pinned, device)
  virtual void map(u64 va, u64 size, u64 handle) = 0;           // cuMemMap
  virtual void set_access(u64 va, u64 size) = 0;                // cuMemSetAccess (RW, device)
  virtual void unmap(u64 va, u64 size) = 0;                     // cuMemUnmap
  virtual void release(u64 handle) = 0;                         // cuMemRelease
  virtual void addr_free(u64 va, u64 size) = 0;                 // cuMemAddressFree
  // make the allocator's context current on the calling thread (mapper thread)
  virtual void bind_thread() = 0;
  // fence support: record an event on `stream` / wait for it on the host.  Two slots: the step that
  // rides behind an in-flight mapper pass records into the slot that pass is not using
  virtual void record_fence(void* stream, int slot = 0) = 0;
  virtual void wait_fence(int slot = 0) = 0;
  virtual bool is_mock() const = 0;
  u64 calls() const { return calls_.load(std::memory_order_relaxed); }

 protected:
  std::atomic<u64> calls_{0};  // the mapper thread and an API thread may both be in the driver
};

class MockVmmDriver : public VmmDriver {
 public:
  explicit MockVmmDriver(u64 granularity = 2ull << 20) : gran_(granularity) {}
  u64 init(int) override { return gran_; }
  u64 reserve(u64 size, u64 alignment) override;
  u64 create(u64 size) override;
  void map(u64 va, u64 size, u64 handle) override;
  void set_access(u64 va, u64 size) override;
  void unmap(u64 va, u64 size) override;
  void release(u64 handle) override;
  void addr_free(u64 va, u64 size) override;
  void bind_thread() override {}
  void record_fence(void*, int slot = 0) override { fence_records_[slot & 1]++; }
  void wait_fence(int slot = 0) override { fence_waits_[slot & 1]++; }
  bool is_mock() const override { return true; }

  std::vector<DriverLogRecord> snapshot_log();
  void clear_log();
  void set_capacity(u64 bytes) { capacity_ = bytes; }  // 0 = unlimited
  // every map / set_access / unmap sleeps this long: a slow driver (eight processes in it at once)
  void set_call_delay_us(u64 us) { delay_us_ = us; }
  u64 fence_records(int slot) const { return fence_records_[slot & 1]; }
  u64 fence_waits(int slot) const { return fence_waits_[slot & 1]; }

 private:
  void log(u64 op, u64 va, u64 size, u64 h);
  u64 gran_;
  u64 next_va_ = 0x7f0000000000ull;  // fake VA space, never dereferenced
  u64 next_handle_ = 1;
  u64 capacity_ = 0, in_use_ = 0;
  std::atomic<u64> delay_us_{0};
  std::atomic<u64> fence_records_[2] = {{0}, {0}}, fence_waits_[2] = {{0}, {0}};
  std::unordered_map<u64, u64> sizes_;
  std::mutex mu_;
  std::vector<DriverLogRecord> log_;
};

// Real driver. Throws std::runtime_error with the driver's error string.
std::unique_ptr<VmmDriver> make_cuda_vmm_driver();

// dlsym'd libcuda entry (shared with the TMA descriptor encoder in the kernels'
// host code).  Returns nullptr when libcuda.so.1 or the symbol is missing.
void* cuda_driver_symbol(const char* name);

}  // namespace vattn
