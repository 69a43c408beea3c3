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

#include "attn_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_desc.h"

namespace vattn {

void launch_combine(const vattn_fwd_params_t& p, int splits, const SplitWorkspace& ws, cudaStream_t stream);

namespace {

using namespace ptx;

constexpr int kTile = 128;                  // keys per tile == TMEM lanes
constexpr int kHeadDim = 128;
constexpr int kTileBytes = kTile * kHeadDim * 2;  // 32 KB
constexpr int kStages = 3;
constexpr int kNPad = 16;                   // MMA N: query heads of the group, zero padded
constexpr int kThreads = 256;
constexpr int kTmemCols = 64;               // S^T x2 (16 cols each) + O^T x2
constexpr int kMaxTilesPerChunk = 16;

struct DecodeTcParams {
  const char* q;
  char* out;
  float* lse;
  float* ws_acc;
  float* ws_ml;
  const int32_t* cache_seqlens;
  const int32_t* cache_batch_idx;
  int64_t q_b, q_h, o_b, o_h;  // byte strides
  int seqlen_k, seqlen_new, num_heads, num_kv_heads, group;
  int tiles_per_chunk, num_chunks;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv;
  uint32_t v_lbo, v_sbo;  // MN-major descriptor strides for the V tile
};

struct __align__(1024) DecodeSmem {
  uint8_t ring[kStages][kTileBytes];     // K / V tiles as TMA wrote them (2 x [128 x 128 B] atoms)
  uint8_t q[2][kNPad * 128];             // Q  : 2 K-atoms of [16 rows x 64 dims], SW128
  uint8_t p[2][2][kNPad * 128];          // P^T: double buffered, 2 K-atoms of [16 rows x 64 keys]
  float wmax[2][4][kNPad];               // cross-warp tile max exchange
  float red[4][kNPad];                   // final row-sum exchange
  uint64_t full[kStages], empty[kStages];
  uint64_t s_full[2], p_ready[2], o_full[2];
  uint32_t tmem_base;
};

// position of V_j / K_j in the load sequence K0, K1, V0, K2, V1, K3, ... of an n-tile chunk
__device__ __forceinline__ int seq_pos_k(int j) { return j < 2 ? j : 2 * j - 1; }
__device__ __forceinline__ int seq_pos_v(int j, int n) {
  const int c0 = n < 2 ? n : 2;
  const int extra = n - 2 > 0 ? (j < n - 2 ? j : n - 2) : 0;
  return c0 + j + extra;
}

// byte offset of 16-bit element (row r, col c) inside a [rows x 64] SW128 K-major atom
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
}

template <typename T, int GP>
__global__ void __launch_bounds__(kThreads, 2)
decode_tc_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                 const DecodeTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  DecodeSmem& sm = *reinterpret_cast<DecodeSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int chunk = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = p.cache_batch_idx ? p.cache_batch_idx[b] : b;
  const int len = (p.cache_seqlens ? p.cache_seqlens[b] : p.seqlen_k) + p.seqlen_new;
  const int ntiles_seq = (len + kTile - 1) / kTile;
  const int tile0 = chunk * p.tiles_per_chunk;
  const int n = min(p.tiles_per_chunk, ntiles_seq - tile0);  // tiles this CTA owns
  const int G = p.group;
  const int h0 = hkv * G;

  if (n <= 0) {
    // nothing to do for this chunk: publish an empty partial so the combine skips it
    if (p.num_chunks > 1 && threadIdx.x < G) {
      const int64_t base = ((int64_t)b * p.num_heads + h0 + threadIdx.x) * p.num_chunks + chunk;
      p.ws_ml[base * 2] = -INFINITY;
      p.ws_ml[base * 2 + 1] = 0.f;
    } else if (p.num_chunks == 1) {
      // zero-length sequence: output zeros (softmax.h:76-78 convention)
      for (int i = threadIdx.x; i < G * kHeadDim; i += kThreads)
        reinterpret_cast<T*>(p.out + b * p.o_b + (int64_t)(h0 + i / kHeadDim) * p.o_h)[i % kHeadDim] =
            Elem<T>::from_f(0.f);
      if (p.lse && threadIdx.x < G) p.lse[(int64_t)b * p.num_heads + h0 + threadIdx.x] = INFINITY;
    }
    return;
  }

  // ---------------------------------------------------------------- setup ----
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
    for (int s = 0; s < kStages; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.p_ready[i], 128);
      mbar_init(&sm.o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&sm.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  {
    // zero Q and P^T (rows >= G must stay zero), then stage this group's query heads
    uint32_t* z = reinterpret_cast<uint32_t*>(sm.q);
    for (int i = threadIdx.x; i < (int)(sizeof(sm.q) + sizeof(sm.p)) / 4; i += kThreads) z[i] = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * (kHeadDim / 8); i += kThreads) {
    const int g = i / (kHeadDim / 8), c8 = i % (kHeadDim / 8);  // 16-byte chunk c8 of head g
    const uint4 v = *reinterpret_cast<const uint4*>(p.q + b * p.q_b + (int64_t)(h0 + g) * p.q_h + c8 * 16);
    const int atom = c8 >> 3, cc = (c8 & 7) * 8;
    *reinterpret_cast<uint4*>(sm.q[atom] + sw128_off(g, cc)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // =========================================================== TMA producer ====
    if (lane == 0) {
      int pos = 0;
      auto load = [&](const CUtensorMap* m, int tile) {
        const int s = pos % kStages;
        mbar_wait(&sm.empty[s], ((pos / kStages) & 1) ^ 1);
        mbar_expect_tx(&sm.full[s], kTileBytes);
        tma_load_5d(sm.ring[s], m, &sm.full[s], 0, (tile0 + tile) * kTile, 0, hkv, slot);
        pos++;
      };
      load(&kmap, 0);
      if (n > 1) load(&kmap, 1);
      for (int j = 0; j < n; j++) {
        load(&vmap, j);
        if (j + 2 < n) load(&kmap, j + 2);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer ====
    if (lane == 0) {
      int pos = 0;
      const uint32_t q_addr = smem_u32(sm.q[0]);
      auto issue_qk = [&](int j) {
        const int s = pos % kStages;
        mbar_wait(&sm.full[s], (pos / kStages) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.ring[s]);
#pragma unroll
        for (int ks = 0; ks < kHeadDim / 16; ks++) {
          const uint32_t a = a0 + (ks >> 2) * (kTile * 128) + (ks & 3) * 32;
          const uint32_t bq = q_addr + (ks >> 2) * (kNPad * 128) + (ks & 3) * 32;
          umma_ss(tmem + (j & 1) * kNPad, make_smem_desc(a, 16, 1024, kLayoutSw128),
                  make_smem_desc(bq, 16, 1024, kLayoutSw128), p.idesc_qk, ks > 0);
        }
        umma_commit(&sm.empty[s]);       // K tile consumed
        umma_commit(&sm.s_full[j & 1]);  // S^T_j ready for the softmax warps
        pos++;
      };
      auto issue_pv = [&](int j) {
        const int s = pos % kStages;
        mbar_wait(&sm.full[s], (pos / kStages) & 1);
        mbar_wait(&sm.p_ready[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.ring[s]);
        const uint32_t p_addr = smem_u32(sm.p[j & 1][0]);
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ks++) {
          const uint32_t a = a0 + ks * (16 * 128);  // 16 keys further down the tile
          const uint32_t bp = p_addr + (ks >> 2) * (kNPad * 128) + (ks & 3) * 32;
          umma_ss(tmem + 2 * kNPad + (j & 1) * kNPad, make_smem_desc(a, p.v_lbo, p.v_sbo, kLayoutSw128),
                  make_smem_desc(bp, 16, 1024, kLayoutSw128), p.idesc_pv, ks > 0);
        }
        umma_commit(&sm.empty[s]);       // V tile consumed
        umma_commit(&sm.o_full[j & 1]);  // O_j^T ready
        pos++;
      };
      issue_qk(0);
      if (n > 1) issue_qk(1);
      for (int j = 0; j < n; j++) {
        issue_pv(j);
        if (j + 2 < n) issue_qk(j + 2);
      }
    }
  } else if (warp >= 4) {
    // ================================================= softmax / accumulate ====
    const int t = threadIdx.x - 128;  // key index inside a tile for S^T, head dim for O^T
    const int sw = warp - 4;          // TMEM lane quadrant of this warp
    const uint32_t lane_base = (uint32_t)(sw * 32) << 16;
    float m_run[GP], l_thr[GP], acc[GP], alpha_prev[GP];
#pragma unroll
    for (int g = 0; g < GP; g++) {
      m_run[g] = -INFINITY;
      l_thr[g] = 0.f;
      acc[g] = 0.f;
      alpha_prev[g] = 1.f;
    }
    auto load_cols = [&](uint32_t col, float (&dst)[GP]) {
      uint32_t r[GP];
      if constexpr (GP == 4) tmem_ld_x4(tmem + lane_base + col, r);
      else if constexpr (GP == 8) tmem_ld_x8(tmem + lane_base + col, r);
      else tmem_ld_x16(tmem + lane_base + col, r);
      tmem_wait_ld();
#pragma unroll
      for (int g = 0; g < GP; g++) dst[g] = __uint_as_float(r[g]);
    };
    auto accumulate_o = [&](int j) {  // acc = acc * alpha_j + O_j^T
      mbar_wait(&sm.o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float o[GP];
      load_cols(2 * kNPad + (j & 1) * kNPad, o);
#pragma unroll
      for (int g = 0; g < GP; g++) acc[g] = fmaf(acc[g], alpha_prev[g], o[g]);
    };

    for (int j = 0; j < n; j++) {
      mbar_wait(&sm.s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float s[GP];
      load_cols((j & 1) * kNPad, s);
      const int key = (tile0 + j) * kTile + t;
      const bool valid = key < len;
      float mx[GP];
#pragma unroll
      for (int g = 0; g < GP; g++) {
        s[g] = valid ? s[g] * p.scale_log2 : -INFINITY;
        mx[g] = s[g];
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
        for (int g = 0; g < GP; g++) mx[g] = fmaxf(mx[g], __shfl_xor_sync(0xffffffffu, mx[g], off));
      // every lane holds every head's warp max after the butterfly; lane g publishes head g
#pragma unroll
      for (int g = 0; g < GP; g++)
        if (lane == g) sm.wmax[j & 1][sw][g] = mx[g];
      named_bar_sync(1, 128);
      // the previous tile's O^T can be folded in while this tile's P is being produced
      if (j > 0) accumulate_o(j - 1);
      const bool tail = (tile0 + j + 1) * kTile > len;  // last, partial tile of the sequence
      uint8_t* pbuf = sm.p[j & 1][t >> 6];
#pragma unroll
      for (int g = 0; g < GP; g++) {
        const float tm = fmaxf(fmaxf(sm.wmax[j & 1][0][g], sm.wmax[j & 1][1][g]),
                               fmaxf(sm.wmax[j & 1][2][g], sm.wmax[j & 1][3][g]));
        const float m_new = fmaxf(m_run[g], tm);  // finite: every tile holds >= 1 valid key
        const float alpha = fast_exp2(m_run[g] - m_new);
        const float pr = fast_exp2(s[g] - m_new);
        m_run[g] = m_new;
        alpha_prev[g] = alpha;
        l_thr[g] = fmaf(l_thr[g], alpha, pr);
        if (g < G) *reinterpret_cast<T*>(pbuf + sw128_off(g, t & 63)) = Elem<T>::from_f(pr);
      }
      if (tail) {
        // rows past the sequence end hold whatever was in memory; P is 0 there, but 0 * NaN
        // would poison O, so blank those V rows in shared memory before the MMA reads them
        const int pv = seq_pos_v(j, n);
        mbar_wait(&sm.full[pv % kStages], (pv / kStages) & 1);
        if (!valid) {
          uint8_t* vt = sm.ring[pv % kStages];
#pragma unroll
          for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 8; c++)
              *reinterpret_cast<uint4*>(vt + a * (kTile * 128) + t * 128 + c * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&sm.p_ready[j & 1]);
    }
    accumulate_o(n - 1);

    // ---- epilogue: row sums across the 128 key-threads, then publish ----
    float lsum[GP];
#pragma unroll
    for (int g = 0; g < GP; g++) lsum[g] = l_thr[g];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
      for (int g = 0; g < GP; g++) lsum[g] += __shfl_xor_sync(0xffffffffu, lsum[g], off);
#pragma unroll
    for (int g = 0; g < GP; g++)
      if (lane == g) sm.red[sw][g] = lsum[g];
    named_bar_sync(1, 128);
#pragma unroll
    for (int g = 0; g < GP; g++) {
      if (g >= G) continue;
      const float L = sm.red[0][g] + sm.red[1][g] + sm.red[2][g] + sm.red[3][g];
      const int h = h0 + g;
      if (p.num_chunks == 1) {
        const float inv = L > 0.f ? 1.f / L : 0.f;
        reinterpret_cast<T*>(p.out + b * p.o_b + (int64_t)h * p.o_h)[t] = Elem<T>::from_f(acc[g] * inv);
        if (p.lse && t == 0)
          p.lse[(int64_t)b * p.num_heads + h] = L > 0.f ? (m_run[g] + log2f(L)) * 0.6931471805599453f : INFINITY;
      } else {
        const int64_t base = ((int64_t)b * p.num_heads + h) * p.num_chunks + chunk;
        p.ws_acc[base * kHeadDim + t] = acc[g];
        if (t == 0) {
          p.ws_ml[base * 2] = m_run[g];
          p.ws_ml[base * 2 + 1] = L;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kTmemCols);
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
  const int64_t total = nt * p.batch * p.num_kv_heads;
  // aim for >= ~8 CTAs per resident slot (148 SMs x 2), chunks of at most 16 tiles (2048 keys)
  int64_t tpc = total / (148 * 2 * 8);
  if (tpc < 1) tpc = 1;
  if (tpc > kMaxTilesPerChunk) tpc = kMaxTilesPerChunk;
  return (int)tpc;
}

int num_chunks_for(const vattn_fwd_params_t& p) {
  const int nt = (p.seqlen_k + kTile - 1) / kTile;
  const int tpc = tiles_per_chunk_for(p);
  const int c = (nt + tpc - 1) / tpc;
  return c < 1 ? 1 : c;
}

template <typename T>
void launch_t(const vattn_fwd_params_t& p, void* ws, cudaStream_t stream) {
  const int group = p.num_heads / p.num_kv_heads;
  const int eb = 2;
  DecodeTcParams dp;
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
  dp.tiles_per_chunk = tiles_per_chunk_for(p);
  dp.num_chunks = num_chunks_for(p);
  dp.scale_log2 = p.softmax_scale * kLog2e;
  const uint32_t fmt = p.dtype == VATTN_DTYPE_BF16 ? kFmtBF16 : kFmtF16;
  dp.idesc_qk = make_idesc(fmt, kTile, kNPad, 0, 0);
  dp.idesc_pv = make_idesc(fmt, kHeadDim, kNPad, 1, 0);
  // MN-major SW128 operand: LBO = distance between the two 64-dim atoms, SBO = distance between
  // 8-key groups.  VATTN_UMMA_MN_VARIANT=1 swaps them (kept switchable for the device self test).
  dp.v_lbo = kTile * 128, dp.v_sbo = 1024;
  if (env_int("VATTN_UMMA_MN_VARIANT", 0) == 1) dp.v_lbo = 1024, dp.v_sbo = kTile * 128;
  SplitWorkspace w{nullptr, nullptr};
  if (dp.num_chunks > 1) w = carve_workspace(ws, p.batch, p.num_heads, dp.num_chunks, kHeadDim);
  dp.ws_acc = w.acc, dp.ws_ml = w.ml;

  const CUtensorMap kmap = make_headdim128_map(p.k_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch,
                                               p.k_row_stride * eb, p.k_head_stride * eb,
                                               p.k_batch_stride * eb, kTile);
  const CUtensorMap vmap = make_headdim128_map(p.v_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch,
                                               p.v_row_stride * eb, p.v_head_stride * eb,
                                               p.v_batch_stride * eb, kTile);
  const size_t smem = sizeof(DecodeSmem) + 1024;
  dim3 grid(dp.num_chunks, p.num_kv_heads, p.batch);
  auto launch = [&](auto kernel) {
    // all GP instantiations share one function-pointer type, so a static flag here would be
    // shared between them; the attribute call is idempotent and cheap, set it every time
    VATTN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tslot = timing_begin(stream);
    kernel<<<grid, kThreads, smem, stream>>>(kmap, vmap, dp);
    timing_end(tslot, stream);
  };
  if (group <= 4) launch(decode_tc_kernel<T, 4>);
  else if (group <= 8) launch(decode_tc_kernel<T, 8>);
  else launch(decode_tc_kernel<T, 16>);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
  if (dp.num_chunks > 1) launch_combine(p, dp.num_chunks, w, stream);
}

}  // namespace

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
  // a 128-row TMA box must never straddle into an unmapped 2 MB page: the row pitch has to
  // divide 16 KB so that tokens_per_page is a multiple of 128 (megacache views fail this)
  const int64_t pitch_k = p.k_row_stride * 2, pitch_v = p.v_row_stride * 2;
  if (pitch_k <= 0 || pitch_v <= 0 || 16384 % pitch_k != 0 || 16384 % pitch_v != 0)
    return no("row pitch does not divide 16 KB (TMA tile could cross an unmapped page)");
  if ((p.k_head_stride * 2) % 16 || (p.k_batch_stride * 2) % 16 || (p.v_head_stride * 2) % 16 ||
      (p.v_batch_stride * 2) % 16)
    return no("strides not 16-byte multiples");
  if (p.cache_batch < 1) return no("cache_batch unknown");
  return true;
}

size_t decode_tc_workspace(const vattn_fwd_params_t& p) {
  const int c = num_chunks_for(p);
  return c > 1 ? split_workspace_bytes(p.batch, p.num_heads, c, kHeadDim) : 0;
}

void launch_decode_tc(const vattn_fwd_params_t& p, void* ws, size_t, cudaStream_t stream) {
  if (p.dtype == VATTN_DTYPE_BF16) launch_t<__nv_bfloat16>(p, ws, stream);
  else launch_t<__half>(p, ws, stream);
}

}  // namespace vattn
