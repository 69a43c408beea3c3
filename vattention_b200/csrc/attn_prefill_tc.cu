// Prefill / chunked-prefill attention on sm_100a tensor cores.
//
// Replaces, for seqlen_q > 1 and head_dim == 128: FA-2's forward kernel the reference dispatches
// to for every prefill chunk (vattention_flashattention_wrapper.py:159-166; flashinfer's
// single_prefill for fi_vattn, vattention_flashinfer_wrapper.py:151-158; arithmetic
// pod_attn/pod_attn/flash_fwd_kernel.h:56-500: S = QK^T :324, mask :333, softmax_rescale_o :348,
// O += PV :372, normalise :438).  Tensor-core bound: 4*Hq*D*(c*p + c(c+1)/2) FLOP per chunk.
//
// One CTA = 128 query rows of one (batch, q head).  Per 128-key tile j:
//   S_j[128 x 128]  = Q[128 x 128] . K_j^T        tcgen05.mma SS, Q and K_j staged by TMA (SW128),
//                                                 fp32 accumulator in TMEM (double buffered)
//   softmax warps: thread i <-> query row i <-> TMEM lane i; row max / exp2 / row sum are
//                  thread-local (no shuffles); P_j packed to 16 bit and stored back to TMEM
//   O[128 x 128]   += P_j . V_j                   tcgen05.mma TS: A = P_j from TMEM, B = V_j tile
//                                                 read N-major (dims contiguous), accumulates in TMEM
// The running max used for scaling is only advanced when a row's new max exceeds it by more than
// 2^8 (lazy rescale): O then needs a TMEM read-modify-write for that warp, otherwise none.
// Tiles above the causal diagonal are skipped; only diagonal / tail tiles are masked.
#include <cstdlib>

#include "attn_common.cuh"
#include "sm100_ptx.cuh"
#include "attn_tc_host.h"
#include "attn_tc_work.cuh"
#include "tma_desc.h"

namespace vattn {

namespace {

using namespace ptx;
using namespace tcwork;

constexpr int kBM = kTile, kBN = kTile, kD = kHeadDim;

struct __align__(1024) PrefillKernelSmem {
  PrefillSmem data;
  TcBarriers bar;
  uint32_t tmem_base;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
prefill_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                  const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap kmap_tail,
                  const __grid_constant__ CUtensorMap vmap_tail, const PrefillParams p) {
  extern __shared__ uint8_t smem_raw[];
  PrefillKernelSmem& sm =
      *reinterpret_cast<PrefillKernelSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && (threadIdx.x & 31) == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
  }
  if (warp == 2) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  // heavy (late) row blocks first: under a causal mask they own the most key tiles
  const int mt = p.num_m_tiles - 1 - blockIdx.x;
  prefill_work<T>(&qmap, &kmap, &vmap, &kmap_tail, &vmap_tail, p, sm.data, sm.bar, tmem, mt, blockIdx.y,
                  blockIdx.z, false);
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// The co-resident POD arrangement (prefill_work<T, 3, true>): Q 32 KB + 3 x 32 KB ring, TMEM S x2 (256
// columns, P in place) + O (128 columns, second allocation) -- leaves 98 KB of shared memory, 128 TMEM
// columns and (at <= 184 registers x 256 threads) a quarter of the register file for a decode CTA.
constexpr int kLeanStages = 3;
struct __align__(1024) PrefillLeanKernelSmem {
  PrefillSmemT<kLeanStages> data;
  TcBarriers bar;
  uint32_t tmem_base, tmem_base_o;
};

// 160 registers x 256 threads + a decode CTA's 96 x 256 = the whole register file
template <typename T>
__global__ void __maxnreg__(160)
prefill_lean_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                       const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap kmap_tail,
                       const __grid_constant__ CUtensorMap vmap_tail, const PrefillParams p) {
  extern __shared__ uint8_t smem_raw[];
  PrefillLeanKernelSmem& sm =
      *reinterpret_cast<PrefillLeanKernelSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && (threadIdx.x & 31) == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
  }
  if (warp == 2) {
    tmem_alloc(&sm.tmem_base, 256);
    tmem_alloc(&sm.tmem_base_o, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base, tmem_o = sm.tmem_base_o;
  const int mt = p.num_m_tiles - 1 - blockIdx.x;
  prefill_work<T, kLeanStages, true>(&qmap, &kmap, &vmap, &kmap_tail, &vmap_tail, p, sm.data, sm.bar, tmem, mt,
                                     blockIdx.y, blockIdx.z, false, tmem_o);
  __syncthreads();
  if (warp == 2) {
    tmem_dealloc(tmem, 256);
    tmem_dealloc(tmem_o, 128);
  }
}

struct __align__(1024) Prefill2KernelSmem {
  Prefill2Smem data;
  TcBarriers bar;
  uint32_t tmem_base;
};

// two 128-row blocks per CTA with one softmax warpgroup each (prefill2_work)
template <typename T, int MODE>
__global__ void __launch_bounds__(kPrefill2Threads, 1)
prefill2_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                   const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap kmap_tail,
                   const __grid_constant__ CUtensorMap vmap_tail, const PrefillParams p) {
  extern __shared__ uint8_t smem_raw[];
  Prefill2KernelSmem& sm =
      *reinterpret_cast<Prefill2KernelSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && (threadIdx.x & 31) == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
  }
  if (warp == 2) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int pairs = (p.num_m_tiles + 1) / 2;
  // REGS: 384 threads x 168 registers at launch; inside prefill2_work the producer / MMA warpgroup
  // drops to 56 and the two softmax warpgroups take 224 each (128 x 56 + 256 x 224 = 384 x 168)
  prefill2_work<T, MODE>(&qmap, &kmap, &vmap, &kmap_tail, &vmap_tail, p, sm.data, sm.bar, tmem,
                         pairs - 1 - blockIdx.x, blockIdx.y, blockIdx.z, false);
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <typename T>
void launch_t(const vattn_fwd_params_t& p, cudaStream_t stream) {
  PrefillTcLaunch L;
  build_prefill_tc(p, &L);
  const int tslot = timing_begin(stream);
  // two row blocks per CTA need enough (pair, head, batch) items to fill the SMs; short chunks
  // keep one block per CTA
  const long long pair_items = (long long)((L.pp.num_m_tiles + 1) / 2) * p.num_heads * p.batch;
  if (t_pod_lean) {
    const size_t smem = sizeof(PrefillLeanKernelSmem) + 1024;
    VATTN_CUDA(cudaFuncSetAttribute(prefill_lean_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(L.pp.num_m_tiles, p.num_heads, p.batch);
    prefill_lean_tc_kernel<T><<<grid, kThreads, smem, stream>>>(L.qmap, L.kmap, L.vmap, L.kmap_tail, L.vmap_tail,
                                                                L.pp);
  } else if (p.seqlen_q > kBM && pair_items >= num_sms() && !env_int("VATTN_PREFILL_SINGLE", 0)) {
    const size_t smem = sizeof(Prefill2KernelSmem) + 1024;
    dim3 grid((L.pp.num_m_tiles + 1) / 2, p.num_heads, p.batch);
    // VATTN_PREFILL_REGS=1: S row read from TMEM once and kept in registers (setmaxnreg); =2: additionally
    // one exponential in four on the FMA pipe (poly_exp2).  Written but not yet measured, hence opt-in
    auto launch2 = [&](auto kernel) {
      VATTN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<grid, kPrefill2Threads, smem, stream>>>(L.qmap, L.kmap, L.vmap, L.kmap_tail, L.vmap_tail, L.pp);
    };
    const int mode = env_int("VATTN_PREFILL_REGS", 0);
    if (mode == 2) launch2(prefill2_tc_kernel<T, 2>);
    else if (mode == 1) launch2(prefill2_tc_kernel<T, 1>);
    else launch2(prefill2_tc_kernel<T, 0>);
  } else {
    const size_t smem = sizeof(PrefillKernelSmem) + 1024;
    VATTN_CUDA(cudaFuncSetAttribute(prefill_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(L.pp.num_m_tiles, p.num_heads, p.batch);
    prefill_tc_kernel<T><<<grid, kThreads, smem, stream>>>(L.qmap, L.kmap, L.vmap, L.kmap_tail, L.vmap_tail,
                                                           L.pp);
  }
  timing_end(tslot, stream);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
}

}  // namespace

void build_prefill_tc(const vattn_fwd_params_t& p, PrefillTcLaunch* out) {
  const int eb = 2;
  PrefillParams& pp = out->pp;
  pp.out = (char*)p.out;
  pp.lse = p.softmax_lse;
  pp.cache_seqlens = p.cache_seqlens;
  pp.cache_batch_idx = p.cache_batch_idx;
  pp.o_b = p.o_batch_stride * eb, pp.o_r = p.o_row_stride * eb, pp.o_h = p.o_head_stride * eb;
  pp.seqlen_q = p.seqlen_q, pp.seqlen_k = p.seqlen_k;
  pp.seqlen_new = p.k_new ? p.seqlen_new : 0;
  pp.num_heads = p.num_heads;
  pp.batch = p.batch;
  pp.group = p.num_heads / p.num_kv_heads;
  pp.num_m_tiles = (p.seqlen_q + kBM - 1) / kBM;
  pp.causal = p.causal;
  pp.scale_log2 = p.softmax_scale * kLog2e;
  const uint32_t fmt = p.dtype == VATTN_DTYPE_BF16 ? kFmtBF16 : kFmtF16;
  pp.idesc_qk = make_idesc(fmt, kBM, kBN, 0, 0);
  pp.idesc_pv = make_idesc(fmt, kBM, kD, 0, 1);  // A = P from TMEM (K-major), B = V tile N-major
  pp.v_lbo = kBN * 128, pp.v_sbo = 1024;
  out->qmap = make_headdim128_map(p.q, p.seqlen_q, p.num_heads, p.batch, p.q_row_stride * eb,
                                  p.q_head_stride * eb, p.q_batch_stride * eb, kBM);
  out->kmap = make_headdim128_map(p.k_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.k_row_stride * eb,
                                  p.k_head_stride * eb, p.k_batch_stride * eb, kBN);
  out->vmap = make_headdim128_map(p.v_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.v_row_stride * eb,
                                  p.v_head_stride * eb, p.v_batch_stride * eb, kBN);
  const int rk = safe_tail_rows(p.k_row_stride * eb), rv = safe_tail_rows(p.v_row_stride * eb);
  pp.tail_rows = rk < rv ? rk : rv;
  out->kmap_tail = make_headdim128_map(p.k_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.k_row_stride * eb,
                                       p.k_head_stride * eb, p.k_batch_stride * eb, pp.tail_rows, 1);
  out->vmap_tail = make_headdim128_map(p.v_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.v_row_stride * eb,
                                       p.v_head_stride * eb, p.v_batch_stride * eb, pp.tail_rows, 1);
}

bool prefill_tc_supported(const vattn_fwd_params_t& p, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (env_int("VATTN_DISABLE_TC", 0)) return no("disabled by VATTN_DISABLE_TC");
  if (p.head_dim != kD) return no("head_dim != 128");
  if (p.seqlen_k < 1) return no("empty cache");
  if (safe_tail_rows(p.k_row_stride * 2) == 0 || safe_tail_rows(p.v_row_stride * 2) == 0)
    return no("row pitch neither divides 16 KB nor divides 2 MB into >= 8-row pages");
  const int64_t st[] = {p.q_row_stride, p.q_head_stride, p.q_batch_stride, p.k_head_stride,
                        p.k_batch_stride, p.v_head_stride, p.v_batch_stride};
  for (int64_t s : st)
    if ((s * 2) % 16) return no("strides not 16-byte multiples");
  if (p.cache_batch < 1) return no("cache_batch unknown");
  return true;
}

size_t prefill_tc_workspace(const vattn_fwd_params_t&) { return 0; }

void launch_prefill_tc(const vattn_fwd_params_t& p, void*, size_t, cudaStream_t stream) {
  if (p.dtype == VATTN_DTYPE_BF16) launch_t<__nv_bfloat16>(p, stream);
  else launch_t<__half>(p, stream);
}

}  // namespace vattn
