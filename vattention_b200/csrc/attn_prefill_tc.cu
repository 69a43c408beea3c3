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
  // MODE 1: 384 threads x 168 registers at launch; the producer / MMA warpgroup drops to 56 and the two
  // softmax warpgroups take 224 each (128 x 56 + 256 x 224 = 384 x 168)
  PrefillSegment seg{};
  seg.mt2 = pairs - 1 - blockIdx.x, seg.h = blockIdx.y, seg.b = blockIdx.z;
  seg.j0 = 0, seg.j1 = INT_MAX, seg.parts = 1;
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

// ---- stream-K persistent prefill ---------------------------------------------------------------
// One CTA per SM.  Items = (row-block pair, q head, batch entry); an item with n visible key tiles
// contributes max(1, n) virtual tiles.  Per (batch entry, kv head) group, the flattened tile space,
// ordered pair (heavy first) -> q head of the group -> key tile, is cut into equal ranges per CTA ON
// THE DEVICE (lengths come from cache_seqlens), so all SMs finish a group together whatever the item
// count (the grid kernel runs 256 items of chunk 2048 as 1.73 waves on 148 SMs; chunk 512 fills only
// 128), all CTAs stream the same kv head's K/V at any moment (L2 resident), and a short chunk deep in
// a long context is split along the keys.  An item cut by a range boundary is reduced by its last
// part to arrive (prefill2_work).  Workspace: 2 partial slots per (group, CTA).
// (First version: equal ranges over the whole launch -- measured 898 vs 986 TFLOP/s for the grid
// kernel at chunk 2048: all four kv heads, 256 MB of K/V, were in flight at once.)
constexpr int kSkMaxEntries = 4096;  // (batch entry, pair) prefix kept in shared memory

struct __align__(1024) PrefillSkSmem {
  Prefill2Smem data;
  TcBarriers bar;
  uint32_t tmem_base;
  int warp_sum[kPrefill2Threads / 32];
  int prefix[kSkMaxEntries + 1];  // exclusive prefix of virtual tiles over (batch entry, pair rank)
};

struct PrefillSkArgs {
  float* ws_o;
  float* ws_ml;
  int* arrive;   // one counter per CTA (the item that starts in that CTA's range and runs past its end)
  int pairs;     // row-block pairs per (batch entry, head)
};

template <typename T, int MODE>
__global__ void __launch_bounds__(kPrefill2Threads, 1)
prefill_sk_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                  const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap kmap_tail,
                  const __grid_constant__ CUtensorMap vmap_tail, const PrefillParams p, const PrefillSkArgs a) {
  extern __shared__ uint8_t smem_raw[];
  PrefillSkSmem& sm =
      *reinterpret_cast<PrefillSkSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs = a.pairs, E = p.batch * pairs;
  auto entry_vt = [&](int e) {  // virtual tiles of (batch entry e / pairs, pair rank e % pairs), heavy first
    const int b = e / pairs, mt2 = pairs - 1 - e % pairs;
    const int lk = (p.cache_seqlens ? p.cache_seqlens[b] : p.seqlen_k) + p.seqlen_new;
    int m0[2], rows[2], nt[2];
    prefill_pair_tiles(p, lk, mt2, m0, rows, nt);
    return max(1, max(nt[0], nt[1]));
  };
  // ---- block-wide exclusive scan of the entries into shared memory
  const int per = (E + kPrefill2Threads - 1) / kPrefill2Threads;
  const int e0 = min(E, (int)threadIdx.x * per), e1 = min(E, e0 + per);
  int mine = 0;
  for (int e = e0; e < e1; e++) mine += entry_vt(e);
  int incl = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) sm.warp_sum[warp] = incl;
  if (threadIdx.x == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
    // every barrier word holds a live mbarrier from here on, so segments can inval + re-init
    for (int s = 0; s < kMaxStages; s++) {
      mbar_init(&sm.bar.full[s], 1);
      mbar_init(&sm.bar.empty[s], 1);
    }
    mbar_init(&sm.bar.q_full, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&sm.bar.s_full[i], 1);
      mbar_init(&sm.bar.p_ready[i], 1);
      mbar_init(&sm.bar.o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    int run = incl - mine;
    for (int w = 0; w < warp; w++) run += sm.warp_sum[w];
    for (int e = e0; e < e1; e++) {
      sm.prefix[e] = run;
      run += entry_vt(e);
    }
    if (e1 == E && (e0 < E || threadIdx.x == 0)) sm.prefix[E] = run;
  }
  __syncthreads();
  const uint32_t tmem = sm.tmem_base;
  const int G = p.group, Hkv = p.num_heads / p.group;
  const int cta = blockIdx.x, ctas = gridDim.x;
  // One (batch entry, kv head) group at a time, its G x pairs items cut into equal ranges per CTA:
  // every SM streams the SAME kv head's K/V at any moment (64 MB at 128K context: L2 resident),
  // whereas equal ranges over the whole launch would put all kv heads in flight at once.
  auto walk = [&](auto role) {
    constexpr int ROLE = decltype(role)::value;
    bool live = true;  // barriers were initialised above
    for (int gi = 0; gi < p.batch * Hkv; gi++) {
      const int b = gi / Hkv, hkv = gi - b * Hkv;
      const int* pre = sm.prefix + b * pairs;  // pre[k] - pre[0]: virtual tiles of pair ranks < k
      StreamKPlan pl;
      pl.total = (int64_t)(pre[pairs] - pre[0]) * G;
      pl.ctas = ctas;
      pl.q = pl.total / ctas;
      pl.r = (int)(pl.total % ctas);
      const int64_t lo = sk_range_begin(pl, cta);
      const int64_t hi = lo + pl.q + (cta < pl.r ? 1 : 0);
      if (lo >= hi) continue;
      // locate tile `lo`: pair rank, q head of the group, key tile
      int k = 0;
      {
        int l = 0, r = pairs;  // largest k with G * (pre[k] - pre[0]) <= lo
        while (r - l > 1) {
          const int m = (l + r) / 2;
          if ((int64_t)(pre[m] - pre[0]) * G <= lo) l = m;
          else r = m;
        }
        k = l;
      }
      int64_t rem = lo - (int64_t)(pre[k] - pre[0]) * G;
      int vt = pre[k + 1] - pre[k];
      int g = (int)(rem / vt);
      int tile = (int)(rem - (int64_t)g * vt);
      for (int64_t x = lo; x < hi;) {
        const int64_t item_start = x - tile, item_end = item_start + vt;
        const int64_t seg_end = hi < item_end ? hi : item_end;
        PrefillSegment seg{};
        seg.mt2 = pairs - 1 - k, seg.h = hkv * G + g, seg.b = b;
        seg.j0 = tile, seg.j1 = tile + (int)(seg_end - x);
        const int first_cta = sk_cta_of(pl, item_start), last_cta = sk_cta_of(pl, item_end - 1);
        seg.parts = last_cta - first_cta + 1;
        seg.ws_o = a.ws_o, seg.ws_ml = a.ws_ml;
        seg.slot_base = (int64_t)gi * ctas * 2;  // every group has its own slots: no reuse hazard
        seg.my_slot = seg.slot_base + 2 * cta + (lo >= item_start ? 0 : 1);
        seg.first_cta = first_cta, seg.item_start = item_start, seg.plan = pl;
        seg.arrive = a.arrive + (int64_t)gi * ctas + first_cta;
        prefill2_work<T, MODE, ROLE>(&qmap, &kmap, &vmap, &kmap_tail, &vmap_tail, p, sm.data, sm.bar, tmem, seg, live);
        x = seg_end;  // a segment ends at hi or at its item's end
        tile = 0;
        if (++g == G) {
          g = 0;
          k++;
          if (k < pairs) vt = pre[k + 1] - pre[k];
        }
      }
    }
  };
  // role split at the top level: each side's code is dominated by its own setmaxnreg
  if (warp < 4) {
    if constexpr (MODE > 0) setmaxnreg_dec<56>();
    walk(std::integral_constant<int, 1>{});
  } else {
    if constexpr (MODE > 0) setmaxnreg_inc<224>();
    walk(std::integral_constant<int, 2>{});
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

int* prefill_arrive_counters(cudaStream_t stream, size_t need) { return arrival_counters(stream, 1, need); }

template <typename T>
void launch_t(const vattn_fwd_params_t& p, void* ws, cudaStream_t stream) {
  PrefillTcLaunch L;
  build_prefill_tc(p, &L);
  const int tslot = timing_begin(stream);
  const int pairs = (L.pp.num_m_tiles + 1) / 2;
  const long long pair_items = (long long)pairs * p.num_heads * p.batch;
  static const bool grid_forced = [] {
    const char* e = std::getenv("VATTN_PREFILL_SCHED");
    return e && std::string(e) == "grid";
  }();
  if (p.seqlen_q > kBM && !grid_forced && (long long)p.batch * pairs <= kSkMaxEntries && ws &&
             prefill_tc_workspace(p) > 0) {
    // stream-K persistent kernel: one CTA per SM, the work split computed on the device
    const int ctas = num_sms();
    PrefillSkArgs a;
    a.pairs = pairs;
    const size_t groups = (size_t)p.batch * p.num_kv_heads;
    a.arrive = prefill_arrive_counters(stream, groups * ctas);
    a.ws_o = static_cast<float*>(ws);
    a.ws_ml = a.ws_o + groups * 2 * ctas * 2 * kBM * kD;
    const size_t smem = sizeof(PrefillSkSmem) + 1024;
    // softmax flavour: 1 = every exponential on the MUFU, 2 = packed pairs + 3/8 of them on the FMA pipe
    static const int mode = env_int("VATTN_PREFILL_MODE", 1);
    auto launch_sk = [&](auto kernel) {
      VATTN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<ctas, kPrefill2Threads, smem, stream>>>(L.qmap, L.kmap, L.vmap, L.kmap_tail, L.vmap_tail, L.pp, a);
    };
    if (mode == 2) launch_sk(prefill_sk_kernel<T, 2>);
    else launch_sk(prefill_sk_kernel<T, 1>);
  } else if (p.seqlen_q > kBM && pair_items >= num_sms() && !env_int("VATTN_PREFILL_SINGLE", 0)) {
    // two row blocks per CTA, one CTA per (pair, head, batch entry): needs enough items to fill the SMs
    const size_t smem = sizeof(Prefill2KernelSmem) + 1024;
    dim3 grid(pairs, p.num_heads, p.batch);
    VATTN_CUDA(cudaFuncSetAttribute(prefill2_tc_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prefill2_tc_kernel<T, 1><<<grid, kPrefill2Threads, smem, stream>>>(L.qmap, L.kmap, L.vmap, L.kmap_tail,
                                                                       L.vmap_tail, L.pp);
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
  // stream-K: two partial slots per (group, CTA), each [2 blocks][128 rows][128 + 2] fp32; beyond
  // 512 MB (many kv heads x batch entries: plenty of items anyway) the grid kernel runs instead
  if (p.seqlen_q <= kBM) return 0;
  const size_t need = (size_t)p.batch * p.num_kv_heads * 2 * num_sms() * 2 * kBM * (kD + 2) * sizeof(float);
  return need <= ((size_t)512 << 20) ? need : 0;
}

void launch_prefill_tc(const vattn_fwd_params_t& p, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (prefill_tc_workspace(p) == 0 || ws_bytes < prefill_tc_workspace(p)) ws = nullptr;
  if (p.dtype == VATTN_DTYPE_BF16) launch_t<__nv_bfloat16>(p, ws, stream);
  else launch_t<__half>(p, ws, stream);
}

}  // namespace vattn
