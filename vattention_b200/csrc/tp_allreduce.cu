// One-shot all-reduce over NVLink peer memory for the head-sharded attention block.
//
// The only collective on the path is the all-reduce(sum) of the row-parallel o_proj output
// [num_tokens, hidden] (sarathi/model_executor/parallel_utils/tensor_parallel/layers.py:448-451 ->
// mappings.py:16-26, NCCL in the reference).  For decode that message is small (64 x 4096 bf16 =
// 512 KB), so the cost is latency, not bandwidth: every rank's partial lives in a symmetric
// (peer-mapped) buffer and each rank sums all of them straight out of its peers' memory with
// 128-bit NVLink loads -- one kernel, no ring steps, no staging copies.
//
// Protocol (per call, epoch e, buffers double-buffered on e & 1):
//   1. the partial of this rank was produced on this stream before the launch (the o_proj GEMM
//      writes directly into the symmetric buffer);
//   2. one thread per peer publishes flag[my_rank] = e into that peer's flag array
//      (system-scope release after a system fence);
//   3. every CTA waits until its own flag array shows e from all ranks (system-scope acquire);
//   4. each thread sums its 16-byte chunks across the `world` peer buffers in fp32 and stores bf16.
// No trailing barrier: a buffer of parity e & 1 is next written by the GEMM of call e + 2, which
// is stream-ordered after this rank's call e + 1, which waited for every peer's flag e + 1, which
// a peer only publishes after its own call e (its reads of our buffer) has completed.
#include "attn_common.cuh"
#include "capi_common.h"

namespace vattn {

namespace {

constexpr int kMaxWorld = 8;
constexpr int kArThreads = 256;

struct ArPeers {
  const uint4* part[kMaxWorld];  // each rank's partial (this epoch's parity)
  uint32_t* flags[kMaxWorld];    // each rank's flag array [world] (this epoch's parity)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data is produced by another GPU: bypass the (non-coherent) L1
__device__ __forceinline__ uint4 ld_peer_128(const uint4* p) {
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kArThreads)
allreduce_oneshot_kernel(const ArPeers peers, uint4* __restrict__ out, int64_t n_chunks, int rank,
                         int world, uint32_t epoch) {
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
  }
  if (threadIdx.x < world) {
    const uint32_t* mine = peers.flags[rank] + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)kArThreads + threadIdx.x; i < n_chunks;
       i += (int64_t)gridDim.x * kArThreads) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < kMaxWorld; r++) {
      if (r >= world) break;
      const uint4 v = ld_peer_128(peers.part[r] + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float2 f = Elem<T>::to_f2(w[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    uint4 o;
    o.x = Elem<T>::from_f2(acc[0], acc[1]);
    o.y = Elem<T>::from_f2(acc[2], acc[3]);
    o.z = Elem<T>::from_f2(acc[4], acc[5]);
    o.w = Elem<T>::from_f2(acc[6], acc[7]);
    out[i] = o;
  }
}

}  // namespace
}  // namespace vattn

using namespace vattn;

extern "C" int vattn_allreduce_oneshot(const uint64_t* peer_partial_ptrs, const uint64_t* peer_flag_ptrs,
                                       void* out, int64_t n_elems, int dtype, int rank, int world,
                                       uint32_t epoch, void* stream) {
  try {
    if (!peer_partial_ptrs || !peer_flag_ptrs || !out) throw ArgError("[vattn] allreduce: null pointer");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
      throw ArgError("[vattn] allreduce: world must be 1..8 and rank inside it");
    if (n_elems % 8 != 0) throw ArgError("[vattn] allreduce: element count must be a multiple of 8");
    if (dtype != VATTN_DTYPE_F16 && dtype != VATTN_DTYPE_BF16) throw ArgError("[vattn] allreduce: fp16/bf16 only");
    ArPeers peers;
    for (int r = 0; r < kMaxWorld; r++) {
      peers.part[r] = r < world ? reinterpret_cast<const uint4*>(peer_partial_ptrs[r]) : nullptr;
      peers.flags[r] = r < world ? reinterpret_cast<uint32_t*>(peer_flag_ptrs[r]) : nullptr;
    }
    const int64_t n_chunks = n_elems / 8;
    int blocks = (int)((n_chunks + kArThreads - 1) / kArThreads);
    if (blocks > 64) blocks = 64;  // all CTAs spin on flags: keep them co-resident by a wide margin
    if (blocks < 1) blocks = 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == VATTN_DTYPE_BF16)
      allreduce_oneshot_kernel<__nv_bfloat16><<<blocks, kArThreads, 0, s>>>(
          peers, static_cast<uint4*>(out), n_chunks, rank, world, epoch);
    else
      allreduce_oneshot_kernel<__half><<<blocks, kArThreads, 0, s>>>(peers, static_cast<uint4*>(out), n_chunks,
                                                                      rank, world, epoch);
    count_launch();
    VATTN_CUDA(cudaGetLastError());
    return VATTN_OK;
  } catch (const ArgError& e) {
    g_last_error = e.what();
    return VATTN_ERR_INVALID;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return VATTN_ERR_DRIVER;
  }
}
