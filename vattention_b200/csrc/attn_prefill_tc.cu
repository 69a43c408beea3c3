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
#include <climits>
#include <type_traits>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>

#include "attn_common.cuh"
#include "sm100_ptx.cuh"
#include "attn_tc_host.h"
#include "attn_tc_work.cuh"
#include "tma_desc.h"

namespace vattn {

size_t prefill_tc_workspace(const vattn_fwd_params_t& p);

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

struct __align__(1024) Prefill2KernelSmem {
  Prefill2Smem data;
  TcBarriers bar;
  uint32_t tmem_base;
};

// Split items (see prefill2_work): arguments of the grid kernel when every item is cut into up to
// `splits` segments along the keys
struct PrefillSplitArgs {
  float* ws_o;   // [item][split][2 blocks][128 rows][128] fp32
  float* ws_ml;  // [item][split][2 blocks][128 rows][2]
  int* arrive;   // one counter per item
  int splits;    // 1: no split (the pointers are unused)
};
constexpr int kMinTilesPerSplit = 16;  // a segment shorter than 16 key tiles is not worth its prologue

// two 128-row blocks per CTA with one softmax warpgroup each (prefill2_work); grid (pairs * splits,
// q heads, batch entries), heavy pairs first, the splits of one item adjacent
template <typename T, int MODE>
__global__ void __launch_bounds__(kPrefill2Threads, 1)
prefill2_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                   const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap kmap_tail,
                   const __grid_constant__ CUtensorMap vmap_tail, const PrefillParams p, const PrefillSplitArgs a) {
  extern __shared__ uint8_t smem_raw[];
  Prefill2KernelSmem& sm =
      *reinterpret_cast<Prefill2KernelSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int pairs = (p.num_m_tiles + 1) / 2;
  PrefillSegment seg{};
  seg.mt2 = pairs - 1 - (int)(blockIdx.x / a.splits), seg.h = blockIdx.y, seg.b = blockIdx.z;
  seg.j0 = 0, seg.j1 = INT_MAX, seg.parts = 1;
  if (a.splits > 1) {
    // the split is decided on the device from the ACTUAL length: parts = min(splits, tiles / 16),
    // identical in every CTA of the item; surplus CTAs leave at once
    const int part = (int)(blockIdx.x % a.splits);
    const int lk = (p.cache_seqlens ? p.cache_seqlens[seg.b] : p.seqlen_k) + p.seqlen_new;
    int m0[2], rows[2], nt[2];
    prefill_pair_tiles(p, lk, seg.mt2, m0, rows, nt);
    const int n = max(nt[0], nt[1]);
    const int parts = max(1, min(a.splits, n / kMinTilesPerSplit));
    if (part >= parts) return;  // CTA-uniform, nothing allocated yet
    const int per = (n + parts - 1) / parts;
    seg.j0 = part * per, seg.j1 = min(n, seg.j0 + per);
    seg.parts = parts;
    const int64_t item = ((int64_t)seg.b * p.num_heads + seg.h) * pairs + (blockIdx.x / a.splits);
    seg.ws_o = a.ws_o, seg.ws_ml = a.ws_ml;
    seg.slot_base = item * a.splits;
    seg.my_slot = seg.slot_base + part;
    seg.arrive = a.arrive + item;
  }
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
  // role split at the top level so that each side's code is dominated by its own setmaxnreg (ptxas
  // allocates registers per setmaxnreg region).  MODE 1: 384 threads x 168 registers at launch; the
  // producer / MMA warpgroup drops to 56 and the two softmax warpgroups take 224 each
  // (128 x 56 + 256 x 224 = 384 x 168)
  if (warp < 4) {
    if constexpr (MODE > 0) setmaxnreg_dec<56>();
    prefill2_work<T, MODE, 1>(&qmap, &kmap, &vmap, &kmap_tail, &vmap_tail, p, sm.data, sm.bar, tmem, seg, false);
  } else {
    if constexpr (MODE > 0) setmaxnreg_inc<224>();
    prefill2_work<T, MODE, 2>(&qmap, &kmap, &vmap, &kmap_tail, &vmap_tail, p, sm.data, sm.bar, tmem, seg, false);
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

int* prefill_arrive_counters(cudaStream_t stream, size_t need) { return arrival_counters(stream, 1, need); }

constexpr int kMaxSplits = 16;
constexpr size_t kSplitSlotBytes = (size_t)2 * kBM * (kD + 2) * sizeof(float);  // one partial: 2 blocks x (O, m, l)
constexpr size_t kSplitWorkspaceCap = (size_t)512 << 20;

// How many segments to cut every (pair, head, batch entry) item into: the smallest s <= 16 whose
// items * s fill the last wave of SMs to >= 95 % (or the best fill found), subject to the workspace
// cap.  256 items (chunk 2048, 32 q heads) -> 4 (1024 units = 6.92 waves); 64 items (chunk 512) -> 16;
// 1024 items (chunk 8192) -> 1.  The device clamps it by the actual number of key tiles.
int prefill_splits(const vattn_fwd_params_t& p) {
  static const int forced = env_int("VATTN_PREFILL_SPLITS", 0);
  if (p.seqlen_q <= kBM) return 1;
  const long long items = (long long)((p.seqlen_q + 2 * kBM - 1) / (2 * kBM)) * p.num_heads * p.batch;
  const int sms = num_sms();
  auto fill = [&](long long units) { return (double)units / (double)(((units + sms - 1) / sms) * sms); };
  int best = 1;
  double best_fill = fill(items);
  for (int s2 = 2; s2 <= kMaxSplits && best_fill < 0.95; s2++) {
    if ((size_t)items * s2 * kSplitSlotBytes > kSplitWorkspaceCap) break;
    const double f = fill(items * s2);
    if (f > best_fill + 0.02) best = s2, best_fill = f;
  }
  if (forced > 0) best = forced > kMaxSplits ? kMaxSplits : forced;
  return best;
}

template <typename T>
void launch_t(const vattn_fwd_params_t& p, void* ws, cudaStream_t stream) {
  PrefillTcLaunch L;
  build_prefill_tc(p, &L);
  const int tslot = timing_begin(stream);
  const int pairs = (L.pp.num_m_tiles + 1) / 2;
  const long long items = (long long)pairs * p.num_heads * p.batch;
  int splits = ws ? prefill_splits(p) : 1;
  // two row blocks per CTA need enough (pair, head, batch) units to fill the SMs; a short chunk with no
  // workspace for the split keeps one block per CTA
  if (p.seqlen_q > kBM && items * splits >= num_sms() && !env_int("VATTN_PREFILL_SINGLE", 0)) {
    const size_t smem = sizeof(Prefill2KernelSmem) + 1024;
    PrefillSplitArgs a{};
    a.splits = splits;
    if (splits > 1) {
      a.arrive = prefill_arrive_counters(stream, (size_t)items);
      a.ws_o = static_cast<float*>(ws);
      a.ws_ml = a.ws_o + (size_t)items * splits * 2 * kBM * kD;
    }
    dim3 grid(pairs * splits, p.num_heads, p.batch);
    VATTN_CUDA(cudaFuncSetAttribute(prefill2_tc_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prefill2_tc_kernel<T, 1><<<grid, kPrefill2Threads, smem, stream>>>(L.qmap, L.kmap, L.vmap, L.kmap_tail,
                                                                       L.vmap_tail, L.pp, a);
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

size_t prefill_tc_workspace(const vattn_fwd_params_t& p) {
  const int splits = prefill_splits(p);
  if (splits <= 1) return 0;
  const size_t items = (size_t)((p.seqlen_q + 2 * kBM - 1) / (2 * kBM)) * p.num_heads * p.batch;
  return items * splits * kSplitSlotBytes;
}

void launch_prefill_tc(const vattn_fwd_params_t& p, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (prefill_tc_workspace(p) == 0 || ws_bytes < prefill_tc_workspace(p)) ws = nullptr;  // no split
  if (p.dtype == VATTN_DTYPE_BF16) launch_t<__nv_bfloat16>(p, ws, stream);
  else launch_t<__half>(p, ws, stream);
}

}  // namespace vattn
