// See vmm_driver.h.  Reference behaviour being replaced: vattention/cudaInternal.h
// (do_cuda_default_init :15-35, reserve_cuda_pages :45-59, map_cuda_pages :70-82,
// do_cuda_kvcache_cleanup :84-94) and vtensor.h:37 (cuMemAddressReserve).
#include "vmm_driver.h"

#include <cuda.h>
#include <dlfcn.h>

#include <chrono>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace vattn {

// ------------------------------------------------------------------ mock ---

void MockVmmDriver::log(u64 op, u64 va, u64 size, u64 h) {
  const u64 us = delay_us_.load(std::memory_order_relaxed);
  if (us && (op == OP_MAP || op == OP_SET_ACCESS || op == OP_UNMAP))
    std::this_thread::sleep_for(std::chrono::microseconds(us));
  std::lock_guard<std::mutex> g(mu_);
  ++calls_;
  log_.push_back({op, va, size, h});
}

u64 MockVmmDriver::reserve(u64 size, u64 alignment) {
  u64 al = alignment ? alignment : gran_;
  u64 va;
  {
    std::lock_guard<std::mutex> g(mu_);
    next_va_ = (next_va_ + al - 1) / al * al;
    va = next_va_;
    next_va_ += size;
  }
  log(OP_RESERVE, va, size, 0);
  return va;
}

u64 MockVmmDriver::create(u64 size) {
  u64 h;
  {
    std::lock_guard<std::mutex> g(mu_);
    if (capacity_ && in_use_ + size > capacity_)
      throw std::runtime_error("[vattn] cuMemCreate failed (mock device out of memory)");
    h = next_handle_++;
    in_use_ += size;
    sizes_[h] = size;
  }
  log(OP_CREATE, 0, size, h);
  return h;
}

void MockVmmDriver::map(u64 va, u64 size, u64 handle) { log(OP_MAP, va, size, handle); }
void MockVmmDriver::set_access(u64 va, u64 size) { log(OP_SET_ACCESS, va, size, 0); }
void MockVmmDriver::unmap(u64 va, u64 size) { log(OP_UNMAP, va, size, 0); }
void MockVmmDriver::release(u64 handle) {
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = sizes_.find(handle);
    if (it != sizes_.end()) {
      in_use_ -= it->second;
      sizes_.erase(it);
    }
  }
  log(OP_RELEASE, 0, 0, handle);
}
void MockVmmDriver::addr_free(u64 va, u64 size) { log(OP_ADDR_FREE, va, size, 0); }

std::vector<DriverLogRecord> MockVmmDriver::snapshot_log() {
  std::lock_guard<std::mutex> g(mu_);
  return log_;
}

void MockVmmDriver::clear_log() {
  std::lock_guard<std::mutex> g(mu_);
  log_.clear();
}

// ------------------------------------------------------------- libcuda ----

namespace {

struct LibCuda {
  void* handle = nullptr;
  using GetProc = CUresult (*)(const char*, void**, int, cuuint64_t);
  GetProc get_proc = nullptr;
  LibCuda() {
    handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (handle) get_proc = reinterpret_cast<GetProc>(dlsym(handle, "cuGetProcAddress"));
  }
  void* sym(const char* name) {
    if (!handle) return nullptr;
    void* fn = nullptr;
    if (get_proc && get_proc(name, &fn, 12000, 0) == CUDA_SUCCESS && fn) return fn;
    return dlsym(handle, name);
  }
};

LibCuda& libcuda() {
  static LibCuda lib;
  return lib;
}

template <typename Fn>
Fn must_sym(const char* name) {
  void* p = libcuda().sym(name);
  if (!p)
    throw std::runtime_error(std::string("[vattn] CUDA driver symbol not available: ") + name +
                             " (is libcuda.so.1 present?)");
  return reinterpret_cast<Fn>(p);
}

class CudaVmmDriver : public VmmDriver {
 public:
  CudaVmmDriver() {
    cuInit_ = must_sym<decltype(cuInit_)>("cuInit");
    cuCtxGetCurrent_ = must_sym<decltype(cuCtxGetCurrent_)>("cuCtxGetCurrent");
    cuCtxSetCurrent_ = must_sym<decltype(cuCtxSetCurrent_)>("cuCtxSetCurrent");
    cuGetErrorString_ = must_sym<decltype(cuGetErrorString_)>("cuGetErrorString");
    cuMemGetAllocationGranularity_ =
        must_sym<decltype(cuMemGetAllocationGranularity_)>("cuMemGetAllocationGranularity");
    cuMemAddressReserve_ = must_sym<decltype(cuMemAddressReserve_)>("cuMemAddressReserve");
    cuMemAddressFree_ = must_sym<decltype(cuMemAddressFree_)>("cuMemAddressFree");
    cuMemCreate_ = must_sym<decltype(cuMemCreate_)>("cuMemCreate");
    cuMemRelease_ = must_sym<decltype(cuMemRelease_)>("cuMemRelease");
    cuMemMap_ = must_sym<decltype(cuMemMap_)>("cuMemMap");
    cuMemUnmap_ = must_sym<decltype(cuMemUnmap_)>("cuMemUnmap");
    cuMemSetAccess_ = must_sym<decltype(cuMemSetAccess_)>("cuMemSetAccess");
    cuEventCreate_ = must_sym<decltype(cuEventCreate_)>("cuEventCreate");
    cuEventRecord_ = must_sym<decltype(cuEventRecord_)>("cuEventRecord");
    cuEventSynchronize_ = must_sym<decltype(cuEventSynchronize_)>("cuEventSynchronize");
    cuEventDestroy_ = must_sym<decltype(cuEventDestroy_)>("cuEventDestroy");
  }

  ~CudaVmmDriver() override {
    for (CUevent e : fence_)
      if (e) cuEventDestroy_(e);
  }

  u64 init(int device) override {
    check(cuInit_(0), "cuInit");
    check(cuCtxGetCurrent_(&ctx_), "cuCtxGetCurrent");
    if (ctx_ == nullptr)
      // the reference prints this and exit(1)s (cudaInternal.h:20-25)
      throw std::runtime_error(
          "[vAttention] No CUDA context found. Please initialize PyTorch before configuring "
          "vAttention.");
    std::memset(&prop_, 0, sizeof(prop_));
    prop_.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop_.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop_.location.id = device;
    std::memset(&access_, 0, sizeof(access_));
    access_.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    access_.location.id = device;
    access_.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t gran = 0;
    check(cuMemGetAllocationGranularity_(&gran, &prop_, CU_MEM_ALLOC_GRANULARITY_MINIMUM),
          "cuMemGetAllocationGranularity");
    return gran;
  }

  u64 reserve(u64 size, u64 alignment) override {
    CUdeviceptr p = 0;
    ++calls_;
    check(cuMemAddressReserve_(&p, size, alignment, 0, 0), "cuMemAddressReserve");
    return static_cast<u64>(p);
  }
  u64 create(u64 size) override {
    CUmemGenericAllocationHandle h = 0;
    ++calls_;
    check(cuMemCreate_(&h, size, &prop_, 0), "cuMemCreate");
    return static_cast<u64>(h);
  }
  void map(u64 va, u64 size, u64 handle) override {
    ++calls_;
    check(cuMemMap_(static_cast<CUdeviceptr>(va), size, 0, handle, 0), "cuMemMap");
  }
  void set_access(u64 va, u64 size) override {
    ++calls_;
    check(cuMemSetAccess_(static_cast<CUdeviceptr>(va), size, &access_, 1), "cuMemSetAccess");
  }
  void unmap(u64 va, u64 size) override {
    ++calls_;
    check(cuMemUnmap_(static_cast<CUdeviceptr>(va), size), "cuMemUnmap");
  }
  void release(u64 handle) override {
    ++calls_;
    check(cuMemRelease_(handle), "cuMemRelease");
  }
  void addr_free(u64 va, u64 size) override {
    ++calls_;
    check(cuMemAddressFree_(static_cast<CUdeviceptr>(va), size), "cuMemAddressFree");
  }
  void bind_thread() override {
    // reference hazard (d) in SURVEY 5: its worker thread never makes the
    // context current; the VMM calls happen to work without one, events don't.
    if (ctx_) check(cuCtxSetCurrent_(ctx_), "cuCtxSetCurrent");
  }
  // a slot is used by one thread at a time: the allocator hands it from the API thread that recorded
  // it to the mapper thread that waits on it under its mutex
  void record_fence(void* stream, int slot = 0) override {
    slot &= 1;
    if (!fence_[slot]) check(cuEventCreate_(&fence_[slot], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    check(cuEventRecord_(fence_[slot], static_cast<CUstream>(stream)), "cuEventRecord");
    fence_armed_[slot] = true;
  }
  void wait_fence(int slot = 0) override {
    slot &= 1;
    if (fence_[slot] && fence_armed_[slot]) {
      check(cuEventSynchronize_(fence_[slot]), "cuEventSynchronize");
      fence_armed_[slot] = false;
    }
  }
  bool is_mock() const override { return false; }

 private:
  void check(CUresult r, const char* what) {
    if (r == CUDA_SUCCESS) return;
    const char* s = nullptr;
    cuGetErrorString_(r, &s);
    throw std::runtime_error(std::string("[vattn] ") + what + " failed (" +
                             std::to_string(static_cast<unsigned>(r)) + "): " + (s ? s : "?"));
  }

  CUcontext ctx_ = nullptr;
  CUmemAllocationProp prop_;
  CUmemAccessDesc access_;
  CUevent fence_[2] = {nullptr, nullptr};
  bool fence_armed_[2] = {false, false};

  CUresult (*cuInit_)(unsigned);
  CUresult (*cuCtxGetCurrent_)(CUcontext*);
  CUresult (*cuCtxSetCurrent_)(CUcontext);
  CUresult (*cuGetErrorString_)(CUresult, const char**);
  CUresult (*cuMemGetAllocationGranularity_)(size_t*, const CUmemAllocationProp*,
                                             CUmemAllocationGranularity_flags);
  CUresult (*cuMemAddressReserve_)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*cuMemAddressFree_)(CUdeviceptr, size_t);
  CUresult (*cuMemCreate_)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                           unsigned long long);
  CUresult (*cuMemRelease_)(CUmemGenericAllocationHandle);
  CUresult (*cuMemMap_)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle,
                        unsigned long long);
  CUresult (*cuMemUnmap_)(CUdeviceptr, size_t);
  CUresult (*cuMemSetAccess_)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*cuEventCreate_)(CUevent*, unsigned);
  CUresult (*cuEventRecord_)(CUevent, CUstream);
  CUresult (*cuEventSynchronize_)(CUevent);
  CUresult (*cuEventDestroy_)(CUevent);
};

}  // namespace

std::unique_ptr<VmmDriver> make_cuda_vmm_driver() {
  return std::unique_ptr<VmmDriver>(new CudaVmmDriver());
}

void* cuda_driver_symbol(const char* name) { return libcuda().sym(name); }

}  // namespace vattn
