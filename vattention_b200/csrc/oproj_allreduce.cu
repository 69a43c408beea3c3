// Row-parallel o_proj GEMM fused with its all-reduce over NVLink peer memory.
//
// Reference: the step right after attention under tensor parallelism is
//     partial = attn_out[tokens, Hq/tp * D] @ W_o_shard^T ;  all_reduce(partial)
// (sarathi/model_executor/parallel_utils/tensor_parallel/layers.py:432-461 -> mappings.py:16-26:
// a cuBLAS GEMM, then NCCL).  Here it is ONE kernel per rank:
//
//   * CTA t owns hidden columns [n t, n t + n), n = 32 for hidden 4096 so that 128 CTAs stream the
//     weight shard (HBM-bound: 2 K hidden bytes, ~1 FLOP/B per token row).  tcgen05: the tokens are
//     the MMA's M axis (<= 128 rows, TMA loads only the real ones), the weight rows its N axis
//     (W_o_shard is [hidden, K] K-major, exactly nn.Linear's layout), the accumulator
//     D[tokens x n] lives in TMEM with lane = token.  One thread streams [n x 64] weight atoms and
//     the [tokens x 64] activation atom (L2 resident) through a 6-stage TMA ring, one thread
//     issues the MMAs.
//   * epilogue (4 warps = 128 TMEM lanes = tokens): accumulator row -> bf16 -> 16-byte stores of
//     the row segment into EVERY rank's receive slot recv[parity][my_rank] over NVLink -> system
//     fence -> per-tile flag on every rank;
//   * the same CTA then waits for the `world` flags of ITS tile, sums the `world` slots (local
//     memory, fixed rank order so every rank gets bit-identical sums) and writes out[tokens, hidden].
//   The tiles are independent, so the exchange of tile t overlaps the weight streaming of the
//   other tiles; no separate collective launch, no partial round trip through HBM.
//
// Epochs live in device memory (epoch_state[4 + tile] = last call this tile completed) so the launch is CUDA-graph
// capturable; the receive slots and flags are double buffered on epoch parity.  Reuse is safe: a
// rank writes parity e & 1 again in call e + 2, after its own call e + 1 finished, in which it
// saw every peer's flag e + 1, which a peer only publishes after ITS call e (the last reader of
// that parity) completed -- launches on one stream run in order.  All CTAs of a call are
// co-resident (grid <= 148, one CTA per SM), so waiting on a peer cannot starve a local tile.
// The exchange moves (world - 1) * tokens * hidden * 2 bytes per rank over NVLink.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "attn_common.cuh"
#include "capi_common.h"
#include "sm100_ptx.cuh"
#include "tma_desc.h"

namespace vattn {

namespace {
using namespace ptx;

constexpr int kMaxWorld = 8;
constexpr int kMaxTileH = 128;  // hidden columns per CTA (MMA N): 32, 64 or 128
constexpr int kKStep = 64;      // K elements per ring stage (one 128-byte swizzle atom)
constexpr int kMaxStages = 16;
constexpr int kRingBytes = 192 * 1024;  // TMA ring; + 16 KB slack the 128-row A descriptor may read into
constexpr int kMaxTokens = 128; // MMA M
constexpr int kThreads = 192;   // warps 0-3 epilogue, 4 TMA, 5 MMA
constexpr uint32_t kSpinLimit = 1u << 24;  // ~20 s of polling before giving up on a peer

struct OprojParams {
  char* recv[kMaxWorld];       // rank r's receive area: [2][world][max_tokens][hidden]
  uint32_t* flags[kMaxWorld];  // rank r's flags: [2][n_tiles][kMaxWorld]
  char* out;                   // [tokens, hidden]
  uint32_t* epoch_state;       // [2] error (spin limit hit); [4 + tile] last epoch completed by this tile's CTA
  int tokens, tokens_pad, hidden, k_steps, max_tokens, rank, world, n_tile, k_rot, stages;
};

// Ring stage = [tokens_pad rows x 128 B] activation atom followed by [n_tile rows x 128 B] weight atom
// (12 KB for 64 tokens x 32 columns => 16 stages in flight per CTA: the k loop is latency bound, so
// depth is what buys bandwidth).  The MMA's A descriptor always spans 128 rows; rows >= tokens_pad
// read whatever follows in the ring and only feed accumulator lanes nobody reads.
struct __align__(1024) OprojSmem {
  uint8_t ring[kRingBytes + kMaxTokens * 128];
  uint64_t full[kMaxStages], empty[kMaxStages], acc_full;
  uint32_t tmem_base, epoch;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_sys_128(const void* p) {  // written by a peer GPU: not through L1
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_sys_128(void* p, const uint4& v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
oproj_allreduce_kernel(const __grid_constant__ CUtensorMap w_map, const __grid_constant__ CUtensorMap x_map,
                       const OprojParams p) {
  extern __shared__ uint8_t raw[];
  OprojSmem& sm = *reinterpret_cast<OprojSmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const uint32_t tmem_cols = (uint32_t)p.n_tile;  // 32 / 64 / 128: a power of two >= 32

  const uint32_t x_bytes = (uint32_t)p.tokens_pad * 128, stage_bytes = x_bytes + (uint32_t)p.n_tile * 128;
  if (threadIdx.x == 0) {
    // every tile keeps its own epoch (all tiles of a launch hold the same value): no cross-CTA
    // atomic and no "last CTA publishes" chain at the end of the kernel
    sm.epoch = *reinterpret_cast<volatile uint32_t*>(p.epoch_state + 4 + tile) + 1;
  }
  if (warp == 5) {
    tmem_alloc(&sm.tmem_base, tmem_cols);
    tmem_relinquish();
  }
  if (warp == 4 && lane == 0) {
    // ---- TMA producer: initialises the ring's barriers itself and starts streaming at once -- the
    // TMEM allocation, the epoch load and the CTA barrier below overlap the first loads' latency
    prefetch_tensormap(&w_map);
    prefetch_tensormap(&x_map);
    for (int s = 0; s < p.stages; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.acc_full, 1);
    fence_mbar_init();
    // every CTA needs the same [tokens x 64] activation atom at step k: start each tile at a
    // different k so the 128 CTAs do not hammer one L2 line set at the same moment
    const int k_first = (int)(((long long)tile * p.k_rot) % p.k_steps);
    const int first = p.stages < p.k_steps ? p.stages : p.k_steps;
    auto issue = [&](int k) {
      const int s = k % p.stages;
      int kk = k + k_first;
      if (kk >= p.k_steps) kk -= p.k_steps;
      mbar_expect_tx(&sm.full[s], stage_bytes);
      uint8_t* st = sm.ring + (size_t)s * stage_bytes;
      tma_load_2d(st, &x_map, &sm.full[s], kk * kKStep, 0);
      tma_load_2d(st + x_bytes, &w_map, &sm.full[s], kk * kKStep, tile * p.n_tile);
    };
    for (int k = 0; k < first; k++) issue(k);  // the whole ring before anybody else is ready
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t epoch = sm.epoch;
  const int parity = epoch & 1;

  if (warp == 4) {
    // ---- TMA producer, steady state (k >= stages): refill a slot as soon as the MMA released it
    if (lane == 0) {
      const int k_first = (int)(((long long)tile * p.k_rot) % p.k_steps);
      for (int k = p.stages; k < p.k_steps; k++) {
        const int s = k % p.stages;
        int kk = k + k_first;
        if (kk >= p.k_steps) kk -= p.k_steps;
        mbar_wait(&sm.empty[s], ((k / p.stages) - 1) & 1);
        mbar_expect_tx(&sm.full[s], stage_bytes);
        uint8_t* st = sm.ring + (size_t)s * stage_bytes;
        tma_load_2d(st, &x_map, &sm.full[s], kk * kKStep, 0);
        tma_load_2d(st + x_bytes, &w_map, &sm.full[s], kk * kKStep, tile * p.n_tile);
      }
    }
  } else if (warp == 5) {
    // ---- MMA issuer: D[128 tokens x n_tile] += X[128 x 64] . W[n_tile x 64]^T
    if (lane == 0) {
      const uint32_t idesc = make_idesc(std::is_same<T, __half>::value ? kFmtF16 : kFmtBF16, kMaxTokens,
                                        (uint32_t)p.n_tile, 0, 0);
      for (int k = 0; k < p.k_steps; k++) {
        const int s = k % p.stages;
        mbar_wait(&sm.full[s], (k / p.stages) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.ring) + (uint32_t)s * stage_bytes, b0 = a0 + x_bytes;
#pragma unroll
        for (int j = 0; j < kKStep / 16; j++)
          umma_ss(tmem, make_smem_desc(a0 + j * 32, 16, 1024, kLayoutSw128),
                  make_smem_desc(b0 + j * 32, 16, 1024, kLayoutSw128), idesc, (k | j) != 0);
        umma_commit(&sm.empty[s]);
      }
      umma_commit(&sm.acc_full);
    }
  } else {
    // ---- epilogue warps 0-3: TMEM lane = token, columns = this tile's hidden columns
    const int t = threadIdx.x;  // 0..127
    const bool live = t < p.tokens;
    mbar_wait(&sm.acc_full, 0);
    tc_fence_after();
    const size_t slot = (size_t)p.max_tokens * p.hidden * 2;
    const size_t row_off = ((size_t)t * p.hidden + (size_t)tile * p.n_tile) * 2;  // this token's segment
    const size_t my_slot = ((size_t)parity * p.world + p.rank) * slot;
    // push: accumulator row -> 16-bit -> every rank's slot [parity][my rank]
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);  // whole warp, also for dead lanes
      tmem_wait_ld();
      if (live) {
        uint4 v0, v1;
        v0.x = Elem<T>::from_f2(__uint_as_float(r[0]), __uint_as_float(r[1]));
        v0.y = Elem<T>::from_f2(__uint_as_float(r[2]), __uint_as_float(r[3]));
        v0.z = Elem<T>::from_f2(__uint_as_float(r[4]), __uint_as_float(r[5]));
        v0.w = Elem<T>::from_f2(__uint_as_float(r[6]), __uint_as_float(r[7]));
        v1.x = Elem<T>::from_f2(__uint_as_float(r[8]), __uint_as_float(r[9]));
        v1.y = Elem<T>::from_f2(__uint_as_float(r[10]), __uint_as_float(r[11]));
        v1.z = Elem<T>::from_f2(__uint_as_float(r[12]), __uint_as_float(r[13]));
        v1.w = Elem<T>::from_f2(__uint_as_float(r[14]), __uint_as_float(r[15]));
        const size_t off = my_slot + row_off + (size_t)c0 * 2;
#pragma unroll
        for (int rk = 0; rk < kMaxWorld; rk++)
          if (rk < p.world) {
            st_sys_128(p.recv[rk] + off, v0);
            st_sys_128(p.recv[rk] + off + 16, v1);
          }
      }
    }
    tc_fence_before();
    // st.release.sys below is cumulative: after this CTA barrier it orders the pushes of all 128
    // threads before the flag, so no per-thread system fence is needed
    named_bar_sync(1, 128);
    const size_t flag_row = ((size_t)parity * gridDim.x + tile) * kMaxWorld;
    if (t < p.world) {
      st_release_sys_u32(p.flags[t] + flag_row + p.rank, epoch);
      const uint32_t* mine = p.flags[p.rank] + flag_row + t;
      uint32_t spins = 0;
      while ((int32_t)(ld_acquire_sys_u32(mine) - epoch) < 0) {
        if (++spins > kSpinLimit) {  // a peer never arrived: fail visibly instead of hanging the GPU
          p.epoch_state[2] = 1;
          break;
        }
        __nanosleep(32);
      }
    }
    named_bar_sync(1, 128);
    // reduce the `world` slots of this tile, rank order, fp32
    if (live) {
      const char* base = p.recv[p.rank] + (size_t)parity * p.world * slot + row_off;
      for (int c0 = 0; c0 < p.n_tile; c0 += 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int rk = 0; rk < kMaxWorld; rk++) {
          if (rk >= p.world) break;
          const uint4 u = ld_sys_128(base + (size_t)rk * slot + (size_t)c0 * 2);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
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
        *reinterpret_cast<uint4*>(p.out + row_off + (size_t)c0 * 2) = o;
      }
    }
  }
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, tmem_cols);
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(p.epoch_state + 4 + tile) = epoch;  // next launch: +1
}

}  // namespace

}  // namespace vattn

using namespace vattn;

extern "C" {

size_t vattn_oproj_allreduce_recv_bytes(int32_t max_tokens, int32_t hidden, int32_t world) {
  return (size_t)2 * world * max_tokens * hidden * 2;
}
size_t vattn_oproj_allreduce_flag_bytes(int32_t hidden) {
  return (size_t)2 * (hidden / 32) * kMaxWorld * sizeof(uint32_t);  // sized for the smallest tile
}

int vattn_oproj_allreduce(const void* x, int64_t x_row_stride, const void* w, void* out, int32_t tokens,
                          int32_t hidden, int32_t k_local, int32_t dtype, int32_t max_tokens,
                          const uint64_t* peer_recv_ptrs, const uint64_t* peer_flag_ptrs,
                          uint32_t* epoch_state, int32_t rank, int32_t world, void* stream) {
  try {
    if (!x || !w || !out || !peer_recv_ptrs || !peer_flag_ptrs || !epoch_state)
      throw ArgError("[vattn] oproj_allreduce: null pointer");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
      throw ArgError("[vattn] oproj_allreduce: world must be 1..8 and rank inside it");
    if (dtype != VATTN_DTYPE_F16 && dtype != VATTN_DTYPE_BF16)
      throw ArgError("[vattn] oproj_allreduce: fp16/bf16 only");
    if (tokens <= 0 || tokens > kMaxTokens || tokens > max_tokens)
      throw UnsupportedError("[vattn] oproj_allreduce: 1..128 tokens per call (decode batches); larger "
                             "batches use the GEMM + all-reduce pair");
    // hidden columns per CTA: the smallest of 32 / 64 / 128 that keeps the grid co-resident
    int n_tile = 0;
    static const int forced_tile = std::getenv("VATTN_OPROJ_NTILE") ? std::atoi(std::getenv("VATTN_OPROJ_NTILE")) : 0;
    static const int k_rot = std::getenv("VATTN_OPROJ_KROT") ? std::atoi(std::getenv("VATTN_OPROJ_KROT")) : 5;
    for (int n : {32, 64, 128})
      if (hidden % n == 0 && hidden / n <= num_sms() && n >= forced_tile) {
        n_tile = n;
        break;
      }
    if (n_tile == 0 || k_local % kKStep != 0 || k_local <= 0)
      throw UnsupportedError("[vattn] oproj_allreduce: hidden must be a multiple of 32 (<= 18944) and "
                             "the local K a multiple of 64");
    if (x_row_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
        (reinterpret_cast<uintptr_t>(out) & 15))
      throw ArgError("[vattn] oproj_allreduce: tensors must be 16-byte aligned");
    OprojParams p{};
    for (int r = 0; r < kMaxWorld; r++) {
      p.recv[r] = r < world ? reinterpret_cast<char*>(peer_recv_ptrs[r]) : nullptr;
      p.flags[r] = r < world ? reinterpret_cast<uint32_t*>(peer_flag_ptrs[r]) : nullptr;
    }
    p.out = static_cast<char*>(out);
    p.epoch_state = epoch_state;
    p.tokens = tokens;
    p.tokens_pad = (tokens + 7) / 8 * 8;  // whole 8-row swizzle groups
    p.n_tile = n_tile;
    p.k_rot = k_rot;
    p.hidden = hidden;
    p.k_steps = k_local / kKStep;
    p.stages = kRingBytes / ((p.tokens_pad + n_tile) * 128);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (p.stages > p.k_steps) p.stages = p.k_steps;
    p.max_tokens = max_tokens;
    p.rank = rank, p.world = world;
    const CUtensorMap w_map = make_kmajor_map(w, hidden, k_local, (int64_t)k_local * 2, n_tile);
    const CUtensorMap x_map = make_kmajor_map(x, tokens, k_local, x_row_stride * 2, p.tokens_pad);
    const size_t smem = sizeof(OprojSmem) + 1024;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == VATTN_DTYPE_BF16) {
      VATTN_CUDA(cudaFuncSetAttribute(oproj_allreduce_kernel<__nv_bfloat16>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      oproj_allreduce_kernel<__nv_bfloat16><<<hidden / n_tile, kThreads, smem, s>>>(w_map, x_map, p);
    } else {
      VATTN_CUDA(cudaFuncSetAttribute(oproj_allreduce_kernel<__half>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      oproj_allreduce_kernel<__half><<<hidden / n_tile, kThreads, smem, s>>>(w_map, x_map, p);
    }
    count_launch();
    VATTN_CUDA(cudaGetLastError());
    return VATTN_OK;
  } catch (const ArgError& e) {
    g_last_error = e.what();
    return VATTN_ERR_INVALID;
  } catch (const UnsupportedError& e) {
    g_last_error = e.what();
    return VATTN_ERR_UNSUPPORTED;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return VATTN_ERR_DRIVER;
  }
}

}  // extern "C"
