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
#include "tma_desc.h"

namespace vattn {

namespace {

using namespace ptx;

constexpr int kBM = 128, kBN = 128, kD = 128;
constexpr int kTileBytes = kBN * kD * 2;  // 32 KB
constexpr int kStages = 5;
constexpr int kThreads = 256;
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;  // TMEM columns: S x2 | O | P x2 (packed)
constexpr float kRescaleThreshold = 8.f;                // log2 domain

struct PrefillParams {
  char* out;
  float* lse;
  const int32_t* cache_seqlens;
  const int32_t* cache_batch_idx;
  int64_t o_b, o_r, o_h;  // byte strides
  int seqlen_q, seqlen_k, seqlen_new, num_heads, group, num_m_tiles;
  int causal;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv;
  uint32_t v_lbo, v_sbo;
};

struct __align__(1024) PrefillSmem {
  uint8_t q[kBM * kD * 2];
  uint8_t ring[kStages][kTileBytes];
  uint64_t q_full, full[kStages], empty[kStages];
  uint64_t s_full[2], p_ready[2], o_done;
  uint32_t tmem_base;
};

__device__ __forceinline__ int seq_pos_v(int j, int n) {
  const int c0 = n < 2 ? n : 2;
  const int extra = n - 2 > 0 ? (j < n - 2 ? j : n - 2) : 0;
  return c0 + j + extra;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
prefill_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                  const __grid_constant__ CUtensorMap vmap, const PrefillParams p) {
  extern __shared__ uint8_t smem_raw[];
  PrefillSmem& sm = *reinterpret_cast<PrefillSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  // heavy (late) row blocks first: under a causal mask they own the most key tiles
  const int mt = p.num_m_tiles - 1 - blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hkv = h / p.group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = p.cache_batch_idx ? p.cache_batch_idx[b] : b;
  const int lk = (p.cache_seqlens ? p.cache_seqlens[b] : p.seqlen_k) + p.seqlen_new;
  const int m0 = mt * kBM;
  const int rows = min(kBM, p.seqlen_q - m0);
  // key j is visible to query row i iff j < lk and (not causal or j <= i + lk - seqlen_q)  (mask.h:172)
  const int shift = lk - p.seqlen_q;
  int kv_end = lk;
  if (p.causal) kv_end = min(lk, m0 + rows + shift);  // exclusive bound for the block's last row
  if (kv_end < 0) kv_end = 0;
  const int n = (kv_end + kBN - 1) / kBN;

  // ---------------------------------------------------------------- setup ----
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
    mbar_init(&sm.q_full, 1);
    for (int s = 0; s < kStages; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.p_ready[i], 128);
    }
    mbar_init(&sm.o_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // =========================================================== TMA producer ====
    if (lane == 0 && n > 0) {
      mbar_expect_tx(&sm.q_full, kBM * kD * 2);
      tma_load_5d(sm.q, &qmap, &sm.q_full, 0, m0, 0, h, b);
      int pos = 0;
      auto load = [&](const CUtensorMap* m, int tile) {
        const int s = pos % kStages;
        mbar_wait(&sm.empty[s], ((pos / kStages) & 1) ^ 1);
        mbar_expect_tx(&sm.full[s], kTileBytes);
        tma_load_5d(sm.ring[s], m, &sm.full[s], 0, tile * kBN, 0, hkv, slot);
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
    if (lane == 0 && n > 0) {
      int pos = 0;
      const uint32_t q_addr = smem_u32(sm.q);
      mbar_wait(&sm.q_full, 0);
      auto issue_qk = [&](int j) {
        const int s = pos % kStages;
        mbar_wait(&sm.full[s], (pos / kStages) & 1);
        tc_fence_after();
        const uint32_t k0 = smem_u32(sm.ring[s]);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ks++) {
          const uint32_t off = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
          umma_ss(tmem + kColS + (j & 1) * kBN, make_smem_desc(q_addr + off, 16, 1024, kLayoutSw128),
                  make_smem_desc(k0 + off, 16, 1024, kLayoutSw128), p.idesc_qk, ks > 0);
        }
        umma_commit(&sm.empty[s]);
        umma_commit(&sm.s_full[j & 1]);
        pos++;
      };
      auto issue_pv = [&](int j) {
        const int s = pos % kStages;
        mbar_wait(&sm.full[s], (pos / kStages) & 1);
        mbar_wait(&sm.p_ready[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v0 = smem_u32(sm.ring[s]);
#pragma unroll
        for (int ks = 0; ks < kBN / 16; ks++) {
          // A: 16 keys = 8 packed TMEM columns of P_j; B: 16 key rows further down the V tile
          umma_ts(tmem + kColO, tmem + kColP + (j & 1) * (kBN / 2) + ks * 8,
                  make_smem_desc(v0 + ks * (16 * 128), p.v_lbo, p.v_sbo, kLayoutSw128), p.idesc_pv,
                  (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&sm.empty[s]);
        umma_commit(&sm.o_done);  // completes phase j
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
    // ==================================================== softmax / epilogue ====
    const int i = threadIdx.x - 128;  // query row inside the block == TMEM lane
    const int sw = warp - 4;
    const uint32_t lane_base = (uint32_t)(sw * 32) << 16;
    const int qi = m0 + i;
    // last visible key (inclusive) for this row; < 0 means the row sees nothing
    int limit = lk - 1;
    if (p.causal) limit = min(limit, qi + shift);
    float m_ref = -INFINITY;  // reference max the stored exponentials are relative to
    float l = 0.f;

    for (int j = 0; j < n; j++) {
      mbar_wait(&sm.s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem + lane_base + kColS + (j & 1) * kBN;
      const int key0 = j * kBN;
      const bool need_mask = key0 + kBN - 1 > limit;  // per-thread; uniform for interior tiles
      // pass 1: row max of the tile
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kBN; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(s_addr + c, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e++) {
          float v = __uint_as_float(r[e]);
          if (need_mask && key0 + c + e > limit) v = -INFINITY;
          mx = fmaxf(mx, v);
        }
      }
      mx *= p.scale_log2;  // scale > 0: max commutes with the scaling
      // lazy rescale: advance the reference only if this row outgrew it by > 2^8
      float alpha = 1.f;
      bool grow = mx > m_ref + kRescaleThreshold;
      if (m_ref == -INFINITY && mx > -INFINITY) grow = true;  // first visible key of the row
      if (grow) {
        alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - mx);
        m_ref = mx;
      }
      const bool any_grow = __any_sync(0xffffffffu, grow) && j > 0;
      if (any_grow) {
        // O holds sum_{t<j} P_t V_t relative to the old reference: wait for PV_{j-1}, then scale
        mbar_wait(&sm.o_done, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
          uint32_t r[32];
          tmem_ld_x32(tmem + lane_base + kColO + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
          tmem_st_x32(tmem + lane_base + kColO + c, r);
        }
        tmem_wait_st();
      }
      l *= alpha;
      // pass 2: exponentials, row sum, pack to 16 bit, store P_j to TMEM
      const float mref_safe = (m_ref == -INFINITY) ? 0.f : m_ref;
      const uint32_t p_addr = tmem + lane_base + kColP + (j & 1) * (kBN / 2);
#pragma unroll
      for (int c = 0; c < kBN; c += 64) {
        uint32_t packed[32];
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          uint32_t r[32];
          tmem_ld_x32(s_addr + c + hh * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float v0 = __uint_as_float(r[e]), v1 = __uint_as_float(r[e + 1]);
            float p0 = fast_exp2(fmaf(v0, p.scale_log2, -mref_safe));
            float p1 = fast_exp2(fmaf(v1, p.scale_log2, -mref_safe));
            if (need_mask) {
              if (key0 + c + hh * 32 + e > limit) p0 = 0.f;
              if (key0 + c + hh * 32 + e + 1 > limit) p1 = 0.f;
            }
            l += p0 + p1;
            packed[hh * 16 + e / 2] = Elem<T>::from_f2(p0, p1);
          }
        }
        tmem_st_x32(p_addr + c / 2, packed);
      }
      tmem_wait_st();
      if ((j + 1) * kBN > lk) {
        // tail tile: key rows past the sequence end are uninitialised memory; P is 0 there but
        // 0 * NaN would poison O, so blank those V rows in shared memory first
        const int pv = seq_pos_v(j, n);
        mbar_wait(&sm.full[pv % kStages], (pv / kStages) & 1);
        if (key0 + i >= lk) {
          uint8_t* vt = sm.ring[pv % kStages];
#pragma unroll
          for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 8; c++)
              *reinterpret_cast<uint4*>(vt + a * (kBN * 128) + i * 128 + c * 16) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(&sm.p_ready[j & 1]);
    }

    // ---- epilogue: O / l -> 16 bit -> global ----
    if (n > 0) {
      mbar_wait(&sm.o_done, (n - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    char* orow = p.out + b * p.o_b + (int64_t)qi * p.o_r + (int64_t)h * p.o_h;
#pragma unroll
    for (int c = 0; c < kD; c += 32) {
      uint32_t r[32];
      if (n > 0) {
        tmem_ld_x32(tmem + lane_base + kColO + c, r);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; e++) r[e] = 0;
      }
      if (i < rows) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          o.x = Elem<T>::from_f2(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv);
          o.y = Elem<T>::from_f2(__uint_as_float(r[e + 2]) * inv, __uint_as_float(r[e + 3]) * inv);
          o.z = Elem<T>::from_f2(__uint_as_float(r[e + 4]) * inv, __uint_as_float(r[e + 5]) * inv);
          o.w = Elem<T>::from_f2(__uint_as_float(r[e + 6]) * inv, __uint_as_float(r[e + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + (c + e) * 2) = o;
        }
      }
    }
    if (p.lse && i < rows)
      p.lse[((int64_t)b * p.num_heads + h) * p.seqlen_q + qi] =
          l > 0.f ? (m_ref + log2f(l)) * 0.6931471805599453f : INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <typename T>
void launch_t(const vattn_fwd_params_t& p, cudaStream_t stream) {
  const int eb = 2;
  PrefillParams pp;
  pp.out = (char*)p.out;
  pp.lse = p.softmax_lse;
  pp.cache_seqlens = p.cache_seqlens;
  pp.cache_batch_idx = p.cache_batch_idx;
  pp.o_b = p.o_batch_stride * eb, pp.o_r = p.o_row_stride * eb, pp.o_h = p.o_head_stride * eb;
  pp.seqlen_q = p.seqlen_q, pp.seqlen_k = p.seqlen_k;
  pp.seqlen_new = p.k_new ? p.seqlen_new : 0;
  pp.num_heads = p.num_heads;
  pp.group = p.num_heads / p.num_kv_heads;
  pp.num_m_tiles = (p.seqlen_q + kBM - 1) / kBM;
  pp.causal = p.causal;
  pp.scale_log2 = p.softmax_scale * kLog2e;
  const uint32_t fmt = p.dtype == VATTN_DTYPE_BF16 ? kFmtBF16 : kFmtF16;
  pp.idesc_qk = make_idesc(fmt, kBM, kBN, 0, 0);
  pp.idesc_pv = make_idesc(fmt, kBM, kD, 0, 1);  // A = P from TMEM (K-major), B = V tile N-major
  pp.v_lbo = kBN * 128, pp.v_sbo = 1024;
  if (env_int("VATTN_UMMA_MN_VARIANT", 0) == 1) pp.v_lbo = 1024, pp.v_sbo = kBN * 128;

  const CUtensorMap qmap = make_headdim128_map(p.q, p.seqlen_q, p.num_heads, p.batch, p.q_row_stride * eb,
                                               p.q_head_stride * eb, p.q_batch_stride * eb, kBM);
  const CUtensorMap kmap = make_headdim128_map(p.k_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch,
                                               p.k_row_stride * eb, p.k_head_stride * eb,
                                               p.k_batch_stride * eb, kBN);
  const CUtensorMap vmap = make_headdim128_map(p.v_cache, p.seqlen_k, p.num_kv_heads, p.cache_batch,
                                               p.v_row_stride * eb, p.v_head_stride * eb,
                                               p.v_batch_stride * eb, kBN);
  const size_t smem = sizeof(PrefillSmem) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    VATTN_CUDA(cudaFuncSetAttribute(prefill_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(pp.num_m_tiles, p.num_heads, p.batch);
  const int tslot = timing_begin(stream);
  prefill_tc_kernel<T><<<grid, kThreads, smem, stream>>>(qmap, kmap, vmap, pp);
  timing_end(tslot, stream);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
}

}  // namespace

bool prefill_tc_supported(const vattn_fwd_params_t& p, std::string* why) {
  auto no = [&](const char* m) {
    if (why) *why = m;
    return false;
  };
  if (env_int("VATTN_DISABLE_TC", 0)) return no("disabled by VATTN_DISABLE_TC");
  if (p.head_dim != kD) return no("head_dim != 128");
  if (p.seqlen_k < 1) return no("empty cache");
  const int64_t pitch_k = p.k_row_stride * 2, pitch_v = p.v_row_stride * 2;
  if (pitch_k <= 0 || pitch_v <= 0 || 16384 % pitch_k != 0 || 16384 % pitch_v != 0)
    return no("row pitch does not divide 16 KB (TMA tile could cross an unmapped page)");
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
