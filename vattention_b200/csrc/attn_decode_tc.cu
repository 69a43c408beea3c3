// Decode attention on sm_100a tensor cores: TMA-staged K/V tiles, tcgen05.mma with the KEYS
// on the MMA M axis, TMEM accumulators, warp-reduce softmax.
//
// Replaces, for seqlen_q == 1, head_dim == 128: FA-2's split-KV decode kernel the reference
// dispatches to (vattention_flashattention_wrapper.py:194-205; arithmetic
// pod_attn/pod_attn/flash_fwd_kernel.h:503-1077, combine :1115+).  HBM-bandwidth bound:
// algorithmic bytes per (batch, kv head) = 2 * 2 B * 128 * len.
//
// Per 128-key tile j of one (batch b, kv head h):
//   S^T[128 keys x 16]   = K_j[128 x 128] . Q^T[128 x 16]     tcgen05.mma, A = K tile (K-major, as
//                                                             TMA wrote it), B = Q (the GQA group's
//                                                             <= 16 query heads, zero padded)
//   softmax warps: thread t <-> key t <-> TMEM lane t; mask, tile max (shuffle + smem), p = exp2
//   O_j^T[128 dims x 16] = V_j^T[128 x 128] . P^T[128 x 16]   tcgen05.mma, A = V tile read MN-major
//                                                             (dims contiguous), B = P^T from smem
//   acc[d][g] = acc[d][g] * alpha_g + O_j^T[d][g]             thread t <-> dim t, fp32 registers
// A 32 KB K tile costs 8 MMAs of 128x16x16 (~64 tensor cycles); the CUDA cores only touch
// 128 x G scores per tile, so the kernel is paced by TMA/HBM, not by instruction issue.
//
// One CTA = (chunk of <= TPC tiles, kv head, batch entry); 2 CTAs per SM, 3 x 32 KB ring each.
// Chunks write un-normalised partials (acc, m, l) that the shared combine kernel reduces.
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

namespace {

using namespace ptx;
using namespace tcwork;

constexpr int kStages = 3;               // 3 x 32 KB ring per CTA, 2 CTAs per SM
constexpr int kMaxTilesPerChunk = 16;


template <int STAGES>
struct __align__(1024) DecodeKernelSmemT {
  DecodeSmemT<STAGES> data;
  TcBarriers bar;
  uint32_t tmem_base;
};
using DecodeKernelSmem = DecodeKernelSmemT<kStages>;

template <typename T, int GP, int STAGES = kStages>
__global__ void __launch_bounds__(kThreads, 2)
decode_tc_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                 const __grid_constant__ CUtensorMap kmap_tail, const __grid_constant__ CUtensorMap vmap_tail,
                 const DecodeTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  DecodeKernelSmemT<STAGES>& sm =
      *reinterpret_cast<DecodeKernelSmemT<STAGES>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && (threadIdx.x & 31) == 0) {
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
  }
  if (warp == 2) {
    tmem_alloc(&sm.tmem_base, kDecodeTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  decode_work<T, GP, STAGES>(&kmap, &vmap, &kmap_tail, &vmap_tail, p, sm.data, sm.bar, tmem, blockIdx.x,
                             blockIdx.y, blockIdx.z, false);
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kDecodeTmemCols);
}


int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

int tiles_per_chunk_for(const vattn_fwd_params_t& p) {
  if (p.num_splits > 0) {
    const int nt = (p.seqlen_k + kTile - 1) / kTile;
    return (nt + p.num_splits - 1) / p.num_splits;
  }
  const int64_t nt = (p.seqlen_k + kTile - 1) / kTile;
  const int64_t seqs = (int64_t)p.batch * p.num_kv_heads;
  const int64_t total = nt * seqs;
  // Large problems: chunks of 16 tiles (2048 keys), many waves, the block scheduler balances them.
  // Small problems (everything fits one wave of 2 CTAs per SM with <= 16 tiles each... or fewer tiles
  // than that): ONE wave -- as many chunks per sequence as resident CTA slots allow, never more CTAs
  // than slots.  The per-CTA prologue (TMEM, barriers, Q, ring fill) and epilogue cost ~8 us, so a
  // partial second wave is expensive: ncu on B16 x Hkv1 x 32K showed 320 CTAs of 13 tiles on 296 slots
  // = 1.08 waves, 68 us, DRAM 49 % busy (profiles/r2_decode_small_grid_ncu.md); measured sweep of
  // round 1: 72.7 us with 16-tile chunks, 91 us with 4, 180 us with 1.
  static const int forced = env_int("VATTN_DECODE_TPC", 0);
  const int64_t slots = (int64_t)num_sms() * 2;
  int64_t tpc = kMaxTilesPerChunk;
  if (total <= slots * kMaxTilesPerChunk) {
    const int64_t cps = seqs >= slots ? 1 : slots / seqs;  // chunks per sequence: seqs * cps <= slots
    tpc = (nt + cps - 1) / cps;
    if (tpc < 1) tpc = 1;
  }
  if (forced > 0) tpc = forced;  // experiments may exceed the cap (the kernel loops over any count)
  return (int)tpc;
}

// (A persistent stream-K schedule -- the flattened (batch, kv head, tile) space cut into equal ranges
// per CTA on the device, partials reduced by the last part to arrive, one launch -- was built and
// measured in round 2 and lost to this grid everywhere: B64 x Hkv8 x 32K 1.216 vs 1.196 ms, B64 x Hkv1
// x 32K 0.174 vs 0.168, B16 x Hkv1 x 32K 0.100 vs 0.083, 128K 0.203 vs 0.189; ncu showed one straggler
// SM at 2x the mean (profiles/r2_decode_small_streamk_ncu_raw.csv).  The block scheduler's dynamic
// balance beats a static equal split; removed.)

int num_chunks_for(const vattn_fwd_params_t& p) {
  const int nt = (p.seqlen_k + kTile - 1) / kTile;
  const int tpc = tiles_per_chunk_for(p);
  const int c = (nt + tpc - 1) / tpc;
  return c < 1 ? 1 : c;
}

template <typename T>
void launch_t(const vattn_fwd_params_t& p, void* ws, cudaStream_t stream) {
  const int group = p.num_heads / p.num_kv_heads;
  DecodeTcLaunch L;
  build_decode_tc(p, ws, stream, &L, true);
  const size_t smem = sizeof(DecodeKernelSmem) + 1024;
  dim3 grid(L.dp.num_chunks, p.num_kv_heads, p.batch);
  auto launch = [&](auto kernel) {
    // all GP instantiations share one function-pointer type, so a static flag here would be
    // shared between them; the attribute call is idempotent and cheap, set it every time
    VATTN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tslot = timing_begin(stream);
    kernel<<<grid, kThreads, smem, stream>>>(L.kmap, L.vmap, L.kmap_tail, L.vmap_tail, L.dp);
    timing_end(tslot, stream);
  };
  if (group <= 4) launch(decode_tc_kernel<T, 4>);
  else if (group <= 8) launch(decode_tc_kernel<T, 8>);
  else launch(decode_tc_kernel<T, 16>);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
  // partials are reduced inside the kernel (last chunk to arrive) unless that was switched off
  if (!L.dp.arrive && L.dp.num_chunks > 1) launch_combine(p, L.dp.num_chunks, L.ws, stream);
}

}  // namespace

// One zero-initialised counter array per stream for the in-kernel split combine ("last chunk of a
// sequence to finish reduces the partials"); the kernel resets what it used, so there is no
// per-call memset.  Per stream because two decodes on different streams may overlap in time.
static int env_int_(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

int* decode_arrive_counters(cudaStream_t stream, size_t need) { return arrival_counters(stream, 0, need); }

bool decode_tc_fuses_append(const vattn_fwd_params_t& p) { return p.k_new != nullptr && p.seqlen_new == 1; }

void build_decode_tc(const vattn_fwd_params_t& p, void* ws, cudaStream_t stream, DecodeTcLaunch* out, bool) {
  const int group = p.num_heads / p.num_kv_heads;
  const int eb = 2;
  DecodeTcParams& dp = out->dp;
  dp.q = (const char*)p.q;
  dp.out = (char*)p.out;
  dp.lse = p.softmax_lse;
  dp.cache_seqlens = p.cache_seqlens;
  dp.cache_batch_idx = p.cache_batch_idx;
  dp.q_b = p.q_batch_stride * eb, dp.q_h = p.q_head_stride * eb;
  dp.o_b = p.o_batch_stride * eb, dp.o_h = p.o_head_stride * eb;
  dp.seqlen_k = p.seqlen_k;
  dp.seqlen_new = p.k_new ? p.seqlen_new : 0;
  dp.num_heads = p.num_heads, dp.num_kv_heads = p.num_kv_heads, dp.group = group;
  dp.batch = p.batch;
  const bool fuse = decode_tc_fuses_append(p);
  dp.k_new = fuse ? (const char*)p.k_new : nullptr;
  dp.v_new = fuse ? (const char*)p.v_new : nullptr;
  dp.kn_b = p.knew_batch_stride * eb, dp.kn_h = p.knew_head_stride * eb;
  dp.vn_b = p.vnew_batch_stride * eb, dp.vn_h = p.vnew_head_stride * eb;
  dp.k_cache = (char*)p.k_cache, dp.v_cache = (char*)p.v_cache;
  dp.kc_b = p.k_batch_stride * eb, dp.kc_r = p.k_row_stride * eb, dp.kc_h = p.k_head_stride * eb;
  dp.vc_b = p.v_batch_stride * eb, dp.vc_r = p.v_row_stride * eb, dp.vc_h = p.v_head_stride * eb;
  // split partials are reduced by combine_kernel; VATTN_DECODE_COMBINE_INKERNEL=1 lets the last chunk of
  // a sequence to arrive do it inside the sweep instead (one launch less, but every CTA then pays a
  // fence + atomic in its epilogue: measured 1.219 vs 1.203 ms kernel time at B64 x 32K, a wash overall)
  static const int inkernel_env = env_int_("VATTN_DECODE_COMBINE_INKERNEL", -1);
  // measured (B16 x Hkv1 x 32K, one wave of 288 CTAs): separate combine kernel 72.8 us, in-kernel 87.0 us
  // -- every CTA pays a fence + atomic in its epilogue; the default stays the combine kernel
  const bool inkernel = inkernel_env > 0;
  dp.arrive = inkernel ? decode_arrive_counters(stream, (size_t)p.batch * p.num_kv_heads) : nullptr;
  dp.tiles_per_chunk = tiles_per_chunk_for(p);
  dp.num_chunks = num_chunks_for(p);
  dp.scale_log2 = p.softmax_scale * kLog2e;
  const uint32_t fmt = p.dtype == VATTN_DTYPE_BF16 ? kFmtBF16 : kFmtF16;
  dp.idesc_qk = make_idesc(fmt, kTile, kNPad, 0, 0);
  dp.idesc_pv = make_idesc(fmt, kHeadDim, kNPad, 1, 0);
  // MN-major SW128 operand: LBO = distance between the two 64-dim atoms, SBO = distance between
  // 8-key groups (verified on device by vattn_selftest_umma).
  dp.v_lbo = kTile * 128, dp.v_sbo = 1024;
  out->ws = SplitWorkspace{nullptr, nullptr};
  if (dp.num_chunks > 1) out->ws = carve_workspace(ws, p.batch, p.num_heads, dp.num_chunks, kHeadDim);
  dp.ws_acc = out->ws.acc, dp.ws_ml = out->ws.ml;
  out->kmap = make_headdim128_map(p.k_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.k_row_stride * eb,
                                  p.k_head_stride * eb, p.k_batch_stride * eb, kTile);
  out->vmap = make_headdim128_map(p.v_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.v_row_stride * eb,
                                  p.v_head_stride * eb, p.v_batch_stride * eb, kTile);
  // tail boxes: rows per box small enough to stay inside a mapped page (tma_desc.h)
  const int rk = safe_tail_rows(p.k_row_stride * eb), rv = safe_tail_rows(p.v_row_stride * eb);
  dp.tail_rows = rk < rv ? rk : rv;
  out->kmap_tail = make_headdim128_map(p.k_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.k_row_stride * eb,
                                       p.k_head_stride * eb, p.k_batch_stride * eb, dp.tail_rows, 1);
  out->vmap_tail = make_headdim128_map(p.v_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch, p.v_row_stride * eb,
                                       p.v_head_stride * eb, p.v_batch_stride * eb, dp.tail_rows, 1);
}

bool decode_tc_supported(const vattn_fwd_params_t& p, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (env_int("VATTN_DISABLE_TC", 0)) return no("disabled by VATTN_DISABLE_TC");
  if (p.seqlen_q != 1) return no("seqlen_q != 1");
  if (p.head_dim != kHeadDim) return no("head_dim != 128");
  if (p.num_heads / p.num_kv_heads > kNPad) return no("GQA group > 16");
  if (p.seqlen_k < 1) return no("empty cache");
  // a TMA box must never reach into an unmapped 2 MB page: full 128-row boxes need the row pitch
  // to divide 16 KB; otherwise (megacache views) the last tile of a sequence is fetched in smaller
  // boxes that divide tokens_per_page (safe_tail_rows)
  if (safe_tail_rows(p.k_row_stride * 2) == 0 || safe_tail_rows(p.v_row_stride * 2) == 0)
    return no("row pitch neither divides 16 KB nor divides 2 MB into >= 8-row pages");
  if ((p.k_head_stride * 2) % 16 || (p.k_batch_stride * 2) % 16 || (p.v_head_stride * 2) % 16 ||
      (p.v_batch_stride * 2) % 16)
    return no("strides not 16-byte multiples");
  if (p.cache_batch < 1) return no("cache_batch unknown");
  return true;
}

size_t decode_tc_workspace_grid(const vattn_fwd_params_t& p) {
  const int c = num_chunks_for(p);
  return c > 1 ? split_workspace_bytes(p.batch, p.num_heads, c, kHeadDim) : 0;
}

size_t decode_tc_workspace(const vattn_fwd_params_t& p) { return decode_tc_workspace_grid(p); }

void launch_decode_tc(const vattn_fwd_params_t& p, void* ws, size_t, cudaStream_t stream) {
  if (p.dtype == VATTN_DTYPE_BF16) launch_t<__nv_bfloat16>(p, ws, stream);
  else launch_t<__half>(p, ws, stream);
}

}  // namespace vattn
