// Shared device/host helpers for the attention kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdexcept>
#include <string>

#include "../../include/vattn_b200.h"

namespace vattn {

extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// optional per-launch timing of the dominant kernel (vattn_kernel_timing)
int timing_begin(cudaStream_t stream);          // returns a slot (or -1 when disabled)
void timing_end(int slot, cudaStream_t stream);

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct UnsupportedError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw CudaError(std::string("[vattn] ") + what + ": " + cudaGetErrorString(e));
}
#define VATTN_CUDA(x) ::vattn::cuda_check((x), #x)

constexpr float kLog2e = 1.4426950408889634f;

// SM count of the current device (cudaDevAttrMultiProcessorCount; 148 on a B200), cached per device
inline int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Zero-initialised arrival counters for the in-kernel reductions of the stream-K kernels ("the last
// part of an item to arrive reduces its partials"; the kernels reset what they used).  One slice of
// a per-device pool per (stream, user): two launches on different streams may overlap in time, and
// the pool is created once, so no allocation happens on a stream that is being captured into a CUDA
// graph (capi_attn.cu).
int* arrival_counters(cudaStream_t stream, int user, size_t need);

// Partial results of a split-KV pass: per (batch, q_row, q_head, split) an
// un-normalised fp32 accumulator of head_dim values plus (running max in the
// log2 domain, running sum).  Layout:
//   acc  [batch*seqlen_q][num_heads][num_splits][head_dim]  fp32
//   ml   [batch*seqlen_q][num_heads][num_splits][2]         fp32 (m, l)
struct SplitWorkspace {
  float* acc;
  float* ml;
};

inline size_t split_workspace_bytes(int64_t rows, int heads, int splits, int head_dim) {
  return static_cast<size_t>(rows) * heads * splits * (head_dim + 2) * sizeof(float);
}
inline SplitWorkspace carve_workspace(void* ws, int64_t rows, int heads, int splits, int head_dim) {
  SplitWorkspace w;
  w.acc = static_cast<float*>(ws);
  w.ml = w.acc + static_cast<size_t>(rows) * heads * splits * head_dim;
  return w;
}

// ---------------------------------------------------------------- device ----

template <typename T>
struct Elem;
template <>
struct Elem<__half> {
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float to_f(__half h) { return __half2float(h); }
  static __device__ __forceinline__ __half from_f(float f) { return __float2half_rn(f); }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    // bf16 -> fp32 is a 16-bit shift: two ALU ops per packed pair
    float2 r;
    r.x = __uint_as_float(u << 16);
    r.y = __uint_as_float(u & 0xffff0000u);
    return r;
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float to_f(__nv_bfloat16 h) { return __bfloat162float(h); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float f) { return __float2bfloat16_rn(f); }
};

// streaming 128-bit load: read-only path, do not allocate in L1 (K/V are touched once)
__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_128(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace vattn
