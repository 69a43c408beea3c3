// Device-side work functions of the tensor-core attention kernels.  One call processes one
// work item with the whole CTA (256 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 4-7
// softmax / accumulate):
//   decode_work   one (chunk of 128-key tiles, kv head, batch entry) of a seqlen_q == 1 problem
//   prefill_work  one (128-row query block, q head, batch entry) of a seqlen_q > 1 problem
// The stand-alone decode / prefill kernels call them once; the POD kernel (attn_pod_tc.cu) calls
// either, item after item, from one persistent CTA per SM.  All mbarriers are (re)initialised at
// the start of every item, so a work function has no state that outlives it except TMEM/SMEM
// ownership, which the caller provides.
//
// Algorithm notes and the reference lines each function replaces are in attn_decode_tc.cu and
// attn_prefill_tc.cu.
#pragma once
#include "attn_common.cuh"
#include "sm100_ptx.cuh"

namespace vattn {
namespace tcwork {

using namespace ptx;

constexpr int kThreads = 256;
constexpr int kTile = 128;     // keys per tile == TMEM lanes
constexpr int kHeadDim = 128;
constexpr int kTileBytes = kTile * kHeadDim * 2;  // 32 KB

// position of V_j in the load sequence K0, K1, V0, K2, V1, K3, ... of an n-tile item
__device__ __forceinline__ int seq_pos_v(int j, int n) {
  const int c0 = n < 2 ? n : 2;
  const int extra = n - 2 > 0 ? (j < n - 2 ? j : n - 2) : 0;
  return c0 + j + extra;
}
// byte offset of 16-bit element (row r, col c) inside a [rows x 64] SW128 K-major atom
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
}
__device__ __forceinline__ void mbar_reinit(uint64_t* bar, uint32_t count, bool was_live) {
  if (was_live) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  mbar_init(bar, count);
}


// Which warps of the CTA run one pipeline, and the barrier they synchronise on.  The stand-alone
// kernels and the persistent POD kernel give a whole CTA to one work item at a time (Side{}: warp 0
// TMA producer, warp 1 MMA issuer, warps 4-7 softmax, __syncthreads).  The dual-role POD kernel runs a
// prefill pipeline and a decode pipeline side by side in ONE CTA: each gets its own warps, its own
// named barrier, its own shared memory and TMEM columns, and pulls its own work items.
struct Side {
  int tma_warp = 0, mma_warp = 1, sm_warp0 = 4;  // softmax warps sm_warp0 .. sm_warp0 + 3
  int bar_id = 0;                                // 0: __syncthreads(); else named barrier of `nthreads`
  int nthreads = 0;                              // threads of the side (0: blockDim.x)
  // linear index of the calling thread inside the side (whole-CTA side: threadIdx.x)
  __device__ __forceinline__ int tid() const {
    if (bar_id == 0) return threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == tma_warp) return lane;
    if (warp == mma_warp) return 32 + lane;
    return 64 + (warp - sm_warp0) * 32 + lane;
  }
  __device__ __forceinline__ int size() const { return bar_id == 0 ? (int)blockDim.x : nthreads; }
  __device__ __forceinline__ void sync() const {
    if (bar_id == 0) __syncthreads();
    else named_bar_sync(bar_id, nthreads);
  }
};

// Stage rows [row0, row0 + 128) of one (head, slot) of a K or V cache into a ring slot.  Rows at or
// beyond `safe_rows` (the sequence length rounded up to `tail_rows`, itself a divisor of
// tokens_per_page) may lie in unmapped virtual memory and are never requested: a tile that is
// entirely safe is one 32 KB box; the last tile of a sequence on a layout whose pages hold fewer
// than 128 rows (megacache views) is fetched as tail_rows-row boxes per 64-column atom.
__device__ __forceinline__ void load_kv_tile(uint8_t* dst, const CUtensorMap* full, const CUtensorMap* tail,
                                             uint64_t* bar, int row0, int head, int slot, int safe_rows,
                                             int tail_rows) {
  if (row0 + kTile <= safe_rows) {
    mbar_expect_tx(bar, kTileBytes);
    tma_load_5d(dst, full, bar, 0, row0, 0, head, slot);
  } else {
    const int nbox = (safe_rows - row0 + tail_rows - 1) / tail_rows;
    mbar_expect_tx(bar, nbox * tail_rows * 128 * 2);
    for (int i = 0; i < nbox; i++)
      for (int a = 0; a < 2; a++)
        tma_load_5d(dst + a * (kTile * 128) + i * tail_rows * 128, tail, bar, 0, row0 + i * tail_rows, a, head,
                    slot);
  }
}

// mbarriers of one CTA; shared by both kinds of work so that a persistent CTA alternating between
// them re-initialises the same, never-overwritten words
constexpr int kMaxStages = 6;
struct TcBarriers {
  uint64_t full[kMaxStages], empty[kMaxStages];  // TMA ring
  uint64_t q_full;                               // prefill: Q block landed
  uint64_t s_full[2], p_ready[2];                // S ready for softmax / P ready for the PV MMA
  uint64_t o_full[2];                            // decode: O_j^T ready; prefill uses [0] as "PV_j done"
  int ticket;                                    // split items: arrival order of this part (prefill stream-K)
};

// ======================================================================== decode ====
constexpr int kNPad = 16;          // MMA N: query heads of the GQA group, zero padded
constexpr int kDecodeTmemCols = 64;  // S^T x2 (16 columns each) + O^T x2

struct DecodeTcParams {
  const char* q;
  char* out;
  float* lse;
  float* ws_acc;
  float* ws_ml;
  const int32_t* cache_seqlens;
  const int32_t* cache_batch_idx;
  int64_t q_b, q_h, o_b, o_h;  // byte strides
  int seqlen_k, seqlen_new, num_heads, num_kv_heads, group, batch;
  int tiles_per_chunk, num_chunks;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv;
  uint32_t v_lbo, v_sbo;  // MN-major descriptor strides for the V tile
  int tail_rows;          // rows per tail TMA box (128 = never needed)
  // fused append of ONE new token per sequence (NULL: none, or appended by a separate kernel):
  // the chunk that owns the end of the sequence folds k_new/v_new in from registers and writes
  // them to cache row cache_seqlens[b]
  const char* k_new;
  const char* v_new;
  int64_t kn_b, kn_h, vn_b, vn_h;              // byte strides of k_new / v_new
  char* k_cache;
  char* v_cache;
  int64_t kc_b, kc_r, kc_h, vc_b, vc_r, vc_h;  // byte strides of the caches
  int* arrive;  // one counter per (batch, kv head): the last chunk to finish combines the partials
};

template <int STAGES>
struct __align__(1024) DecodeSmemT {
  uint8_t ring[STAGES][kTileBytes];  // K / V tiles as TMA wrote them (2 x [128 x 128 B] atoms)
  uint8_t q[2][kNPad * 128];         // Q  : 2 K-atoms of [16 rows x 64 dims], SW128
  uint8_t p[2][2][kNPad * 128];      // P^T: double buffered, 2 K-atoms of [16 rows x 64 keys]
  float wmax[2][4][kNPad];           // cross-warp tile max exchange
  float red[4][kNPad];               // final row-sum exchange
  float red2[4][kNPad];              // q . k_new exchange (fused append)
  int ticket;                        // arrival order of this chunk among its sequence's chunks
};

// One contiguous run of key tiles of one (batch entry, kv head): what a CTA processes between two
// (re)initialisations of its pipeline.  The classic grid makes one segment per (chunk, kv head, batch)
// CTA.
struct DecodeSegment {
  int b, hkv;
  int tile0, n;        // first key tile, number of key tiles (0: only the appended token, or nothing)
  int len;             // cached rows of the sequence (without the token appended by this launch)
  bool owns_new;       // this segment folds in / writes the appended token
  int parts;           // segments the sequence is split into (1: the result is final)
  // partial (acc, m, l) of head g of this segment goes to index part_idx + g * part_stride_g;
  // the reducer reads part c of head g at red_idx + g * red_stride_g + c * red_stride_c
  int64_t part_idx, part_stride_g, red_idx, red_stride_g, red_stride_c;
  bool publish_empty;  // nothing to do, but the separate combine kernel expects a slot: write (-inf, 0)
};

template <typename T, int GP, int STAGES>
__device__ void decode_segment(const CUtensorMap* kmap, const CUtensorMap* vmap, const CUtensorMap* kmap_tail,
                               const CUtensorMap* vmap_tail, const DecodeTcParams& p,
                               DecodeSmemT<STAGES>& sm, TcBarriers& bar, uint32_t tmem, const DecodeSegment& seg,
                               bool barriers_live, const Side sd = Side{}) {
  static_assert(STAGES <= kMaxStages, "ring deeper than the barrier block");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = seg.b, hkv = seg.hkv;
  const int slot = p.cache_batch_idx ? p.cache_batch_idx[b] : b;
  const int len = seg.len;
  const int tile0 = seg.tile0;
  const int n = seg.n;
  const bool owns_new = seg.owns_new;
  const int G = p.group;
  const int h0 = hkv * G;

  if (seg.publish_empty) {  // CTA-uniform: this chunk lies past the sequence
    if (sd.tid() < G) {
      const int64_t base = seg.part_idx + sd.tid() * seg.part_stride_g;
      p.ws_ml[base * 2] = -INFINITY;
      p.ws_ml[base * 2 + 1] = 0.f;
    }
    return;
  }
  if (n == 0 && !owns_new) {
    // zero-length sequence, nothing appended: output zeros (softmax.h:76-78 convention)
    for (int i = sd.tid(); i < G * kHeadDim; i += sd.size())
      reinterpret_cast<T*>(p.out + b * p.o_b + (int64_t)(h0 + i / kHeadDim) * p.o_h)[i % kHeadDim] =
          Elem<T>::from_f(0.f);
    if (p.lse && sd.tid() < G) p.lse[(int64_t)b * p.num_heads + h0 + sd.tid()] = INFINITY;
    return;
  }

  // ---------------------------------------------------------------- setup ----
  if (sd.tid() == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_reinit(&bar.full[s], 1, barriers_live);
      mbar_reinit(&bar.empty[s], 1, barriers_live);
    }
    for (int i = 0; i < 2; i++) {
      mbar_reinit(&bar.s_full[i], 1, barriers_live);
      mbar_reinit(&bar.p_ready[i], 128, barriers_live);
      mbar_reinit(&bar.o_full[i], 1, barriers_live);
    }
    fence_mbar_init();
  }
  {
    // zero Q and P^T (rows >= G must stay zero), then stage this group's query heads
    uint32_t* z = reinterpret_cast<uint32_t*>(sm.q);
    for (int i = sd.tid(); i < (int)(sizeof(sm.q) + sizeof(sm.p)) / 4; i += sd.size()) z[i] = 0;
  }
  sd.sync();
  for (int i = sd.tid(); i < G * (kHeadDim / 8); i += sd.size()) {
    const int g = i / (kHeadDim / 8), c8 = i % (kHeadDim / 8);  // 16-byte chunk c8 of head g
    const uint4 v = *reinterpret_cast<const uint4*>(p.q + b * p.q_b + (int64_t)(h0 + g) * p.q_h + c8 * 16);
    *reinterpret_cast<uint4*>(sm.q[c8 >> 3] + sw128_off(g, (c8 & 7) * 8)) = v;
  }
  fence_proxy_async_smem();
  sd.sync();

  if (warp == sd.tma_warp) {
    // =========================================================== TMA producer ====
    if (lane == 0) {
      int pos = 0;
      const int safe_rows = (len + p.tail_rows - 1) / p.tail_rows * p.tail_rows;
      auto load = [&](const CUtensorMap* m, const CUtensorMap* mt, int tile) {
        const int s = pos % STAGES;
        mbar_wait(&bar.empty[s], ((pos / STAGES) & 1) ^ 1);
        load_kv_tile(sm.ring[s], m, mt, &bar.full[s], (tile0 + tile) * kTile, hkv, slot, safe_rows, p.tail_rows);
        pos++;
      };
      if (n > 0) load(kmap, kmap_tail, 0);
      if (n > 1) load(kmap, kmap_tail, 1);
      for (int j = 0; j < n; j++) {
        load(vmap, vmap_tail, j);
        if (j + 2 < n) load(kmap, kmap_tail, j + 2);
      }
    }
  } else if (warp == sd.mma_warp) {
    // ============================================================ MMA issuer ====
    if (lane == 0 && n > 0) {
      int pos = 0;
      const uint32_t q_addr = smem_u32(sm.q[0]);
      auto issue_qk = [&](int j) {
        const int s = pos % STAGES;
        mbar_wait(&bar.full[s], (pos / STAGES) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.ring[s]);
#pragma unroll
        for (int ks = 0; ks < kHeadDim / 16; ks++) {
          const uint32_t a = a0 + (ks >> 2) * (kTile * 128) + (ks & 3) * 32;
          const uint32_t bq = q_addr + (ks >> 2) * (kNPad * 128) + (ks & 3) * 32;
          umma_ss(tmem + (j & 1) * kNPad, make_smem_desc(a, 16, 1024, kLayoutSw128),
                  make_smem_desc(bq, 16, 1024, kLayoutSw128), p.idesc_qk, ks > 0);
        }
        umma_commit(&bar.empty[s]);       // K tile consumed
        umma_commit(&bar.s_full[j & 1]);  // S^T_j ready for the softmax warps
        pos++;
      };
      auto issue_pv = [&](int j) {
        const int s = pos % STAGES;
        mbar_wait(&bar.full[s], (pos / STAGES) & 1);
        mbar_wait(&bar.p_ready[j & 1], (j >> 1) & 1);
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
        umma_commit(&bar.empty[s]);       // V tile consumed
        umma_commit(&bar.o_full[j & 1]);  // O_j^T ready
        pos++;
      };
      issue_qk(0);
      if (n > 1) issue_qk(1);
      for (int j = 0; j < n; j++) {
        issue_pv(j);
        if (j + 2 < n) issue_qk(j + 2);
      }
    }
  } else if (warp >= sd.sm_warp0 && warp < sd.sm_warp0 + 4) {  // (other warps of the side idle here)
    // ================================================= softmax / accumulate ====
    const int t = threadIdx.x - sd.sm_warp0 * 32;  // key index inside a tile for S^T, head dim for O^T
    const int sw = warp - sd.sm_warp0;             // TMEM lane quadrant of this warp (sm_warp0 % 4 == 0)
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
      mbar_wait(&bar.o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float o[GP];
      load_cols(2 * kNPad + (j & 1) * kNPad, o);
#pragma unroll
      for (int g = 0; g < GP; g++) acc[g] = fmaf(acc[g], alpha_prev[g], o[g]);
    };

    for (int j = 0; j < n; j++) {
      mbar_wait(&bar.s_full[j & 1], (j >> 1) & 1);
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
        mbar_wait(&bar.full[pv % STAGES], (pv / STAGES) & 1);
        if (!valid) {
          uint8_t* vt = sm.ring[pv % STAGES];
#pragma unroll
          for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 8; c++)
              *reinterpret_cast<uint4*>(vt + a * (kTile * 128) + t * 128 + c * 16) = make_uint4(0, 0, 0, 0);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&bar.p_ready[j & 1]);
    }
    if (n > 0) accumulate_o(n - 1);

    // ---- epilogue: row sums across the 128 key-threads ----
    float lsum[GP];
#pragma unroll
    for (int g = 0; g < GP; g++) lsum[g] = l_thr[g];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
      for (int g = 0; g < GP; g++) lsum[g] += __shfl_xor_sync(0xffffffffu, lsum[g], off);
    float sn[GP];  // q . k_new partial sums over this thread's dim (fused append)
    float vnew = 0.f;
    if (owns_new) {
      const float kd = Elem<T>::to_f(reinterpret_cast<const T*>(p.k_new + b * p.kn_b + (int64_t)hkv * p.kn_h)[t]);
      vnew = Elem<T>::to_f(reinterpret_cast<const T*>(p.v_new + b * p.vn_b + (int64_t)hkv * p.vn_h)[t]);
#pragma unroll
      for (int g = 0; g < GP; g++)
        sn[g] = g < G ? kd * Elem<T>::to_f(*reinterpret_cast<const T*>(sm.q[t >> 6] + sw128_off(g, t & 63))) : 0.f;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
        for (int g = 0; g < GP; g++) sn[g] += __shfl_xor_sync(0xffffffffu, sn[g], off);
      // the new row goes into the cache for the following steps (nobody reads it in this launch)
      const int64_t row = len;
      if (t < 16)
        *reinterpret_cast<uint4*>(p.k_cache + slot * p.kc_b + row * p.kc_r + (int64_t)hkv * p.kc_h + t * 16) =
            *reinterpret_cast<const uint4*>(p.k_new + b * p.kn_b + (int64_t)hkv * p.kn_h + t * 16);
      else if (t < 32)
        *reinterpret_cast<uint4*>(p.v_cache + slot * p.vc_b + row * p.vc_r + (int64_t)hkv * p.vc_h + (t - 16) * 16) =
            *reinterpret_cast<const uint4*>(p.v_new + b * p.vn_b + (int64_t)hkv * p.vn_h + (t - 16) * 16);
    }
#pragma unroll
    for (int g = 0; g < GP; g++)
      if (lane == g) {
        sm.red[sw][g] = lsum[g];
        if (owns_new) sm.red2[sw][g] = sn[g];
      }
    named_bar_sync(1, 128);
    float Lg[GP];
#pragma unroll
    for (int g = 0; g < GP; g++) {
      Lg[g] = sm.red[0][g] + sm.red[1][g] + sm.red[2][g] + sm.red[3][g];
      if (owns_new) {
        // the appended token as one more key: s = scale * q . k_new, value v_new
        const float s_new = (sm.red2[0][g] + sm.red2[1][g] + sm.red2[2][g] + sm.red2[3][g]) * p.scale_log2;
        const float m_fin = fmaxf(m_run[g], s_new);
        const float a = fast_exp2(m_run[g] - m_fin);  // m_run == -inf (no cached key) -> 0
        const float pn = fast_exp2(s_new - m_fin);
        acc[g] = fmaf(acc[g], a, pn * vnew);
        Lg[g] = fmaf(Lg[g], a, pn);
        m_run[g] = m_fin;
      }
    }

    // one segment covers the whole sequence: final result.  Otherwise publish the partial; it is
    // reduced either by the last segment of the sequence to arrive (p.arrive) or by combine_kernel.
    bool write_out = seg.parts == 1;
    if (!write_out) {
#pragma unroll
      for (int g = 0; g < GP; g++) {
        if (g >= G) continue;
        const int64_t base = seg.part_idx + g * seg.part_stride_g;
        p.ws_acc[base * kHeadDim + t] = acc[g];
        if (t == 0) {
          p.ws_ml[base * 2] = m_run[g];
          p.ws_ml[base * 2 + 1] = Lg[g];
        }
      }
      if (p.arrive) {
        // CTA barrier, then ONE thread fences and takes a ticket (fences are cumulative: the
        // partial stores of the other 127 threads are ordered before the atomic)
        named_bar_sync(1, 128);
        if (t == 0) {
          __threadfence();
          sm.ticket = atomicAdd(p.arrive + (int64_t)b * p.num_kv_heads + hkv, 1);
        }
        named_bar_sync(1, 128);
      }
      if (p.arrive && sm.ticket == seg.parts - 1) {
        __threadfence();
        write_out = true;
        const int np = seg.parts;
        // The ring is idle here (every staged tile was consumed before the last O^T was read): stage
        // all (head, part) statistics in it with ONE round of independent loads, so the reduction is
        // not a chain of dependent L2 round trips (19 parts x 8 heads at B16 x 32K: ~45 us as a chain)
        float* st_m = reinterpret_cast<float*>(sm.ring[0]);
        float* st_l = st_m + G * np;
        const bool staged = (size_t)2 * G * np * sizeof(float) <= (size_t)STAGES * kTileBytes;
        if (staged) {
          for (int idx = t; idx < G * np; idx += 128) {
            const int g = idx / np, c = idx - g * np;
            const int64_t base = seg.red_idx + g * seg.red_stride_g + c * seg.red_stride_c;
            st_m[idx] = __ldcg(p.ws_ml + base * 2);
            st_l[idx] = __ldcg(p.ws_ml + base * 2 + 1);
          }
          named_bar_sync(1, 128);
        }
#pragma unroll
        for (int g = 0; g < GP; g++) {
          if (g >= G) continue;
          const int64_t base = seg.red_idx + g * seg.red_stride_g;
          float M = -INFINITY;
          if (staged) {
            for (int c = 0; c < np; c++) M = fmaxf(M, st_m[g * np + c]);
          } else {
            for (int c = 0; c < np; c++) M = fmaxf(M, __ldcg(p.ws_ml + (base + c * seg.red_stride_c) * 2));
          }
          float o = 0.f, L = 0.f;
#pragma unroll 4
          for (int c = 0; c < np; c++) {
            const int64_t idx = base + c * seg.red_stride_c;
            const float lc = staged ? st_l[g * np + c] : __ldcg(p.ws_ml + idx * 2 + 1);
            const float mc = staged ? st_m[g * np + c] : __ldcg(p.ws_ml + idx * 2);
            // an empty part (l == 0) never wrote its accumulator row: do not read it (0 * garbage)
            const float w = lc > 0.f ? fast_exp2(mc - M) : 0.f;
            const float a = lc > 0.f ? __ldcg(p.ws_acc + idx * kHeadDim + t) : 0.f;
            L = fmaf(lc, w, L);
            o = fmaf(a, w, o);
          }
          acc[g] = o, Lg[g] = L, m_run[g] = M;
        }
        if (t == 0) p.arrive[(int64_t)b * p.num_kv_heads + hkv] = 0;  // ready for the next launch
      }
    }
    if (write_out) {
#pragma unroll
      for (int g = 0; g < GP; g++) {
        if (g >= G) continue;
        const int h = h0 + g;
        const float inv = Lg[g] > 0.f ? 1.f / Lg[g] : 0.f;
        reinterpret_cast<T*>(p.out + b * p.o_b + (int64_t)h * p.o_h)[t] = Elem<T>::from_f(acc[g] * inv);
        if (p.lse && t == 0)
          p.lse[(int64_t)b * p.num_heads + h] =
              Lg[g] > 0.f ? (m_run[g] + log2f(Lg[g])) * 0.6931471805599453f : INFINITY;
      }
    }
  }
  tc_fence_before();
  sd.sync();
}

// Classic grid: one segment per (chunk of tiles_per_chunk tiles, kv head, batch entry); partial layout
// [batch][q head][chunk] as combine_kernel reads it.
template <typename T, int GP, int STAGES>
__device__ void decode_work(const CUtensorMap* kmap, const CUtensorMap* vmap, const CUtensorMap* kmap_tail,
                            const CUtensorMap* vmap_tail, const DecodeTcParams& p,
                            DecodeSmemT<STAGES>& sm, TcBarriers& bar, uint32_t tmem, int chunk, int hkv,
                            int b, bool barriers_live, const Side sd = Side{}) {
  const bool fused_new = p.k_new != nullptr;
  DecodeSegment seg;
  seg.b = b, seg.hkv = hkv;
  // rows read from the cache; with a fused append the new token is NOT read back from memory
  seg.len = (p.cache_seqlens ? p.cache_seqlens[b] : p.seqlen_k) + (fused_new ? 0 : p.seqlen_new);
  const int ntiles_seq = (seg.len + kTile - 1) / kTile;
  const int chunks_active = max(1, (ntiles_seq + p.tiles_per_chunk - 1) / p.tiles_per_chunk);
  seg.tile0 = chunk * p.tiles_per_chunk;
  seg.n = max(0, min(p.tiles_per_chunk, ntiles_seq - seg.tile0));
  seg.owns_new = fused_new && chunk == chunks_active - 1;
  const int64_t row = ((int64_t)b * p.num_heads + hkv * p.group) * p.num_chunks;
  seg.part_idx = row + chunk, seg.part_stride_g = p.num_chunks;
  seg.red_idx = row, seg.red_stride_g = p.num_chunks, seg.red_stride_c = 1;
  // with the in-kernel reduction only the active chunks count; the combine kernel reads every slot
  seg.parts = p.arrive ? chunks_active : p.num_chunks;
  seg.publish_empty = false;
  if (chunk >= chunks_active) {
    if (p.arrive) return;
    seg.publish_empty = true;
  }
  // an empty sequence with nothing appended: with the separate combine kernel its slot 0 must be a
  // well-formed empty partial too (the combine then writes the zeros / +inf lse)
  if (seg.n == 0 && !seg.owns_new && !p.arrive && p.num_chunks > 1) seg.publish_empty = true;
  decode_segment<T, GP, STAGES>(kmap, vmap, kmap_tail, vmap_tail, p, sm, bar, tmem, seg, barriers_live, sd);
}

// ======================================================================= prefill ====
constexpr int kPrefillStages = 5;
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;  // TMEM columns: S x2 | O | P x2 (packed)
constexpr float kRescaleThreshold = 8.f;                // log2 domain

struct PrefillParams {
  char* out;
  float* lse;
  const int32_t* cache_seqlens;
  const int32_t* cache_batch_idx;
  int64_t o_b, o_r, o_h;  // byte strides
  int seqlen_q, seqlen_k, seqlen_new, num_heads, group, num_m_tiles, batch;
  int causal;
  float scale_log2;
  uint32_t idesc_qk, idesc_pv;
  uint32_t v_lbo, v_sbo;
  int tail_rows;
};

template <int STAGES>
struct __align__(1024) PrefillSmemT {
  uint8_t q[kTile * kHeadDim * 2];
  uint8_t ring[STAGES][kTileBytes];
};
using PrefillSmem = PrefillSmemT<kPrefillStages>;

// Row max of one S tile (thread = query row).  MASK is a warp-uniform choice: interior tiles
// (the vast majority) pay no per-element compare.  Four independent max chains: with one softmax
// warp per scheduler there is nothing else to hide the 4-cycle dependent-issue latency.
template <bool MASK>
__device__ __forceinline__ float tile_row_max(uint32_t s_addr, int key0, int limit) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < kTile; c += 32) {
    uint32_t r[32];
    tmem_ld_x32(s_addr + c, r);
    tmem_wait_ld();
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      float v0 = __uint_as_float(r[e]), v1 = __uint_as_float(r[e + 1]);
      float v2 = __uint_as_float(r[e + 2]), v3 = __uint_as_float(r[e + 3]);
      if (MASK) {
        const int k = key0 + c + e;
        if (k > limit) v0 = -INFINITY;
        if (k + 1 > limit) v1 = -INFINITY;
        if (k + 2 > limit) v2 = -INFINITY;
        if (k + 3 > limit) v3 = -INFINITY;
      }
      m0 = fmaxf(m0, v0), m1 = fmaxf(m1, v1), m2 = fmaxf(m2, v2), m3 = fmaxf(m3, v3);
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// p = exp2(s * scale - mref) for one S tile, packed to 16 bit and stored to TMEM as the A operand of
// the PV MMA (2 keys per 32-bit column); returns the row sum of the (unrounded) exponentials.
template <typename T, bool MASK>
__device__ __forceinline__ float tile_exp_store(uint32_t s_addr, uint32_t p_addr, float scale_log2,
                                                float mref, int key0, int limit) {
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
  for (int c = 0; c < kTile; c += 64) {
    uint32_t packed[32];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      uint32_t r[32];
      tmem_ld_x32(s_addr + c + hh * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        float p0 = fast_exp2(fmaf(__uint_as_float(r[e]), scale_log2, -mref));
        float p1 = fast_exp2(fmaf(__uint_as_float(r[e + 1]), scale_log2, -mref));
        float p2 = fast_exp2(fmaf(__uint_as_float(r[e + 2]), scale_log2, -mref));
        float p3 = fast_exp2(fmaf(__uint_as_float(r[e + 3]), scale_log2, -mref));
        if (MASK) {
          const int k = key0 + c + hh * 32 + e;
          if (k > limit) p0 = 0.f;
          if (k + 1 > limit) p1 = 0.f;
          if (k + 2 > limit) p2 = 0.f;
          if (k + 3 > limit) p3 = 0.f;
        }
        l0 += p0, l1 += p1, l2 += p2, l3 += p3;
        packed[hh * 16 + e / 2] = Elem<T>::from_f2(p0, p1);
        packed[hh * 16 + e / 2 + 1] = Elem<T>::from_f2(p2, p3);
      }
    }
    tmem_st_x32(p_addr + c / 2, packed);
  }
  return (l0 + l1) + (l2 + l3);
}

// ---- register-resident S row (prefill2_work<T, true>) ------------------------------------------
// With one softmax warp per scheduler nothing hides the TMEM read latency, and the two-pass scheme
// above reads every S column twice (max pass, exp pass) with a wait after each 32-column load.
// This variant reads the 128-column row ONCE -- four loads in flight, one wait -- and keeps it in
// registers for both passes; it needs ~224 registers per softmax thread, which the kernel gets by
// moving registers from the producer / MMA warpgroup with setmaxnreg.
template <bool MASK>
__device__ __forceinline__ float regs_quarter_max(const uint32_t (&r)[32], int key_base, int limit) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    float v0 = __uint_as_float(r[e]), v1 = __uint_as_float(r[e + 1]);
    float v2 = __uint_as_float(r[e + 2]), v3 = __uint_as_float(r[e + 3]);
    if (MASK) {
      const int k = key_base + e;
      if (k > limit) v0 = -INFINITY;
      if (k + 1 > limit) v1 = -INFINITY;
      if (k + 2 > limit) v2 = -INFINITY;
      if (k + 3 > limit) v3 = -INFINITY;
    }
    m0 = fmaxf(m0, v0), m1 = fmaxf(m1, v1), m2 = fmaxf(m2, v2), m3 = fmaxf(m3, v3);
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// exponentials of one 32-column quarter, packed to 16 bit into out[0..15]; adds to the four partial sums.
// (Sending a fraction of them through a polynomial on the FMA pipe -- scalar, then packed FFMA2 pairs --
// was measured twice and lost both times: 890 vs 974 and 879 vs 901 TFLOP/s at chunk 2048; the MUFU is
// 52-55 % busy, it is the dependency chain S -> softmax -> P -> PV -> QK^T that paces the kernel.)
template <typename T, bool MASK>
__device__ __forceinline__ void regs_quarter_exp(const uint32_t (&r)[32], uint32_t* out, float scale_log2,
                                                 float mref, int key_base, int limit, float (&l)[4]) {
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    float p0 = fast_exp2(fmaf(__uint_as_float(r[e]), scale_log2, -mref));
    float p1 = fast_exp2(fmaf(__uint_as_float(r[e + 1]), scale_log2, -mref));
    float p2 = fast_exp2(fmaf(__uint_as_float(r[e + 2]), scale_log2, -mref));
    const float x3 = fmaf(__uint_as_float(r[e + 3]), scale_log2, -mref);
    float p3 = fast_exp2(x3);
    if (MASK) {
      const int k = key_base + e;
      if (k > limit) p0 = 0.f;
      if (k + 1 > limit) p1 = 0.f;
      if (k + 2 > limit) p2 = 0.f;
      if (k + 3 > limit) p3 = 0.f;
    }
    l[0] += p0, l[1] += p1, l[2] += p2, l[3] += p3;
    out[e / 2] = Elem<T>::from_f2(p0, p1);
    out[e / 2 + 1] = Elem<T>::from_f2(p2, p3);
  }
}

// setmaxnreg (sm_90+): every warp of a warpgroup executes the same instruction
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// LEAN (the co-resident POD arrangement): P_j is written in place over the consumed half of S_j (the
// next writer of that S buffer, QK^T of tile j+2, is issued after PV_j and the tensor core runs in
// order), O lives in a second TMEM allocation `tmem_o`, and the ring is STAGES deep -- 384 TMEM
// columns and 128 KB of shared memory instead of 512 / 192 KB, so that a decode CTA fits beside it.
template <typename T, int STAGES = kPrefillStages, bool LEAN = false>
__device__ void prefill_work(const CUtensorMap* qmap, const CUtensorMap* kmap, const CUtensorMap* vmap,
                             const CUtensorMap* kmap_tail, const CUtensorMap* vmap_tail,
                             const PrefillParams& p, PrefillSmemT<STAGES>& sm, TcBarriers& bar, uint32_t tmem,
                             int mt, int h, int b, bool barriers_live, uint32_t tmem_o = 0,
                             const Side sd = Side{}) {
  constexpr int kStages = STAGES;
  constexpr int kBM = kTile, kBN = kTile, kD = kHeadDim;
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

  if (sd.tid() == 0) {
    mbar_reinit(&bar.q_full, 1, barriers_live);
    for (int s = 0; s < kStages; s++) {
      mbar_reinit(&bar.full[s], 1, barriers_live);
      mbar_reinit(&bar.empty[s], 1, barriers_live);
    }
    for (int i = 0; i < 2; i++) {
      mbar_reinit(&bar.s_full[i], 1, barriers_live);
      mbar_reinit(&bar.p_ready[i], 128, barriers_live);
      mbar_reinit(&bar.o_full[i], 1, barriers_live);
    }
    fence_mbar_init();
  }
  sd.sync();

  if (warp == sd.tma_warp) {
    // =========================================================== TMA producer ====
    if (lane == 0 && n > 0) {
      mbar_expect_tx(&bar.q_full, kBM * kD * 2);
      tma_load_5d(sm.q, qmap, &bar.q_full, 0, m0, 0, h, b);
      int pos = 0;
      const int safe_rows = (lk + p.tail_rows - 1) / p.tail_rows * p.tail_rows;
      auto load = [&](const CUtensorMap* m, const CUtensorMap* mt, int tile) {
        const int s = pos % kStages;
        mbar_wait(&bar.empty[s], ((pos / kStages) & 1) ^ 1);
        load_kv_tile(sm.ring[s], m, mt, &bar.full[s], tile * kBN, hkv, slot, safe_rows, p.tail_rows);
        pos++;
      };
      load(kmap, kmap_tail, 0);
      if (n > 1) load(kmap, kmap_tail, 1);
      for (int j = 0; j < n; j++) {
        load(vmap, vmap_tail, j);
        if (j + 2 < n) load(kmap, kmap_tail, j + 2);
      }
    }
  } else if (warp == sd.mma_warp) {
    // ============================================================ MMA issuer ====
    if (lane == 0 && n > 0) {
      int pos = 0;
      const uint32_t q_addr = smem_u32(sm.q);
      mbar_wait(&bar.q_full, 0);
      auto issue_qk = [&](int j) {
        const int s = pos % kStages;
        mbar_wait(&bar.full[s], (pos / kStages) & 1);
        tc_fence_after();
        const uint32_t k0 = smem_u32(sm.ring[s]);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ks++) {
          const uint32_t off = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
          umma_ss(tmem + kColS + (j & 1) * kBN, make_smem_desc(q_addr + off, 16, 1024, kLayoutSw128),
                  make_smem_desc(k0 + off, 16, 1024, kLayoutSw128), p.idesc_qk, ks > 0);
        }
        umma_commit(&bar.empty[s]);
        umma_commit(&bar.s_full[j & 1]);
        pos++;
      };
      auto issue_pv = [&](int j) {
        const int s = pos % kStages;
        mbar_wait(&bar.full[s], (pos / kStages) & 1);
        mbar_wait(&bar.p_ready[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v0 = smem_u32(sm.ring[s]);
#pragma unroll
        for (int ks = 0; ks < kBN / 16; ks++) {
          // A: 16 keys = 8 packed TMEM columns of P_j; B: 16 key rows further down the V tile
          umma_ts(LEAN ? tmem_o : tmem + kColO,
                  LEAN ? tmem + kColS + (j & 1) * kBN + ks * 8 : tmem + kColP + (j & 1) * (kBN / 2) + ks * 8,
                  make_smem_desc(v0 + ks * (16 * 128), p.v_lbo, p.v_sbo, kLayoutSw128), p.idesc_pv,
                  (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&bar.empty[s]);
        umma_commit(&bar.o_full[0]);  // completes phase j
        pos++;
      };
      issue_qk(0);
      if (n > 1) issue_qk(1);
      for (int j = 0; j < n; j++) {
        issue_pv(j);
        if (j + 2 < n) issue_qk(j + 2);
      }
    }
  } else if (warp >= sd.sm_warp0 && warp < sd.sm_warp0 + 4) {
    // ==================================================== softmax / epilogue ====
    const int i = threadIdx.x - sd.sm_warp0 * 32;  // query row inside the block == TMEM lane
    const int sw = warp - sd.sm_warp0;
    const uint32_t lane_base = (uint32_t)(sw * 32) << 16;
    const int qi = m0 + i;
    // last visible key (inclusive) for this row; < 0 means the row sees nothing
    int limit = lk - 1;
    if (p.causal) limit = min(limit, qi + shift);
    float m_ref = -INFINITY;  // reference max the stored exponentials are relative to
    float l = 0.f;

    for (int j = 0; j < n; j++) {
      mbar_wait(&bar.s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem + lane_base + kColS + (j & 1) * kBN;
      const int key0 = j * kBN;
      const bool need_mask = key0 + kBN - 1 > limit;  // per-thread; false for interior tiles
      // masking is decided per warp so that interior tiles run the compare-free variant
      const bool warp_mask = __any_sync(0xffffffffu, need_mask);
      // pass 1: row max of the tile
      float mx = warp_mask ? tile_row_max<true>(s_addr, key0, limit) : tile_row_max<false>(s_addr, key0, limit);
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
        mbar_wait(&bar.o_full[0], (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
          uint32_t r[32];
          tmem_ld_x32((LEAN ? tmem_o + lane_base + c : tmem + lane_base + kColO + c), r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
          tmem_st_x32((LEAN ? tmem_o + lane_base + c : tmem + lane_base + kColO + c), r);
        }
        tmem_wait_st();
      }
      l *= alpha;
      // pass 2: exponentials, row sum, pack to 16 bit, store P_j to TMEM
      const float mref_safe = (m_ref == -INFINITY) ? 0.f : m_ref;
      const uint32_t p_addr = LEAN ? s_addr : tmem + lane_base + kColP + (j & 1) * (kBN / 2);
      l += warp_mask ? tile_exp_store<T, true>(s_addr, p_addr, p.scale_log2, mref_safe, key0, limit)
                     : tile_exp_store<T, false>(s_addr, p_addr, p.scale_log2, mref_safe, key0, limit);
      tmem_wait_st();
      if ((j + 1) * kBN > lk) {
        // tail tile: key rows past the sequence end are uninitialised memory; P is 0 there but
        // 0 * NaN would poison O, so blank those V rows in shared memory first
        const int pv = seq_pos_v(j, n);
        mbar_wait(&bar.full[pv % kStages], (pv / kStages) & 1);
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
      mbar_arrive(&bar.p_ready[j & 1]);
    }

    // ---- epilogue: O / l -> 16 bit -> global ----
    if (n > 0) {
      // S_{n-1} being ready only proves PV_{n-3} retired, so the barrier may still be in phase
      // n-2: a parity wait for phase n-1 alone would be satisfied by the stale phase n-3 (same
      // parity).  Wait for the two outstanding phases in order.
      if (n > 1) mbar_wait(&bar.o_full[0], (n - 2) & 1);
      mbar_wait(&bar.o_full[0], (n - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    char* orow = p.out + b * p.o_b + (int64_t)qi * p.o_r + (int64_t)h * p.o_h;
#pragma unroll
    for (int c = 0; c < kD; c += 32) {
      uint32_t r[32];
      if (n > 0) {
        tmem_ld_x32((LEAN ? tmem_o + lane_base + c : tmem + lane_base + kColO + c), r);
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
  sd.sync();
}


// ============================================================== prefill, 2 row blocks ====
// Two 128-row query blocks of the same (batch, head) per CTA, each with its own softmax
// warpgroup, sharing every K/V tile (FlashAttention-4's ping-pong): while one warpgroup is in its
// exponentials the tensor core runs the other block's QK^T / PV, so the softmax latency of one
// block is hidden behind the MMAs of the other, and each staged K/V tile feeds twice the FLOPs.
// TMEM: S0 | S1 | O0 | O1 (128 columns each); P_t is written in place over the first 64 columns of
// S_t (16-bit packed) once they have been consumed, and the next QK^T into S_t is issued after
// PV_t in program order (tcgen05.mma executes in issue order).
//
// One call handles key tiles [j0, j1) of one item (row-block pair, q head, batch entry).  When the
// launch has too few items to fill the SMs evenly (256 items are 1.73 waves on 148 SMs; a 512-token
// chunk has 64) the grid kernel splits every item into `parts` segments along the keys: each
// publishes its un-normalised (O, m, l) and the last one to arrive reduces them -- the role of FA-2's
// split-KV kernel + combine (flash_fwd_kernel.h:503-1077,1115+), which the reference needs for short
// chunks deep in a long context (flash_api.cpp:258-323).
constexpr int kPrefill2Threads = 384;
constexpr int kPrefill2Stages = 4;
constexpr uint32_t kCol2S = 0, kCol2O = 256;  // S_t at kCol2S + 128 t, O_t at kCol2O + 128 t

struct __align__(1024) Prefill2Smem {
  uint8_t q[2][kTile * kHeadDim * 2];
  uint8_t ring[kPrefill2Stages][kTileBytes];
};

struct PrefillSegment {
  int mt2, h, b;       // row-block pair, q head, batch entry
  int j0, j1;          // key tiles [j0, j1) of the item (clamped to what each block can see)
  int parts;           // segments the item is split into; 1: this call writes the final result
  // ---- parts > 1 only
  float* ws_o;         // [slot][2 blocks][128 rows][128] fp32, un-normalised
  float* ws_ml;        // [slot][2 blocks][128 rows][2]   (reference max in the log2 domain, row sum)
  int64_t slot_base;   // slot of part c = slot_base + c (the grid kernel's split items)
  int64_t my_slot;
  int* arrive;         // arrival counter of this item (zero before and after the launch)
};

// number of key tiles row block (2 mt2 + t) can see: nt[t]; rows[t] valid query rows; m0[t] first row
__device__ __forceinline__ void prefill_pair_tiles(const PrefillParams& p, int lk, int mt2, int (&m0)[2],
                                                   int (&rows)[2], int (&nt)[2]) {
  const int shift = lk - p.seqlen_q;
#pragma unroll
  for (int t = 0; t < 2; t++) {
    m0[t] = (mt2 * 2 + t) * kTile;
    rows[t] = min(kTile, p.seqlen_q - m0[t]);
    int kv_end = lk;
    if (p.causal) kv_end = min(lk, m0[t] + rows[t] + shift);
    if (kv_end < 0 || rows[t] <= 0) kv_end = 0;
    nt[t] = (kv_end + kTile - 1) / kTile;
  }
}

// MODE 0: the S row is read from TMEM twice (max pass, exp pass).  MODE 1: read once and kept in
// registers for both passes (~224 registers per softmax thread).
// ROLE 0: the whole CTA calls the function (POD kernel).  ROLE 1 / 2: the caller has split the roles
// at the top level -- warps 0-3 call <ROLE 1> (TMA producer, MMA issuer), warps 4-11 call <ROLE 2>
// (softmax) -- so that each side's code is dominated by its own setmaxnreg (ptxas allocates
// registers per setmaxnreg region) and a persistent loop executes setmaxnreg once.  The CTA-wide
// barriers are explicit bar.sync 3, 384, which both sides reach from their own code.
__device__ __forceinline__ void cta_sync_384() { named_bar_sync(3, kPrefill2Threads); }

// Tail of a softmax warpgroup's work on one row block: O / l -> 16 bit -> global when the segment is
// the whole item, otherwise publish the un-normalised partial and let the last part reduce them all.
template <typename T>
__device__ __forceinline__ void prefill_block_epilogue(const PrefillParams& p, const PrefillSegment& seg,
                                                       TcBarriers& bar, uint32_t, uint32_t o_addr, int t, int i,
                                                       int qi, int h, int b, int rows_t, int my_n, float m_ref,
                                                       float l) {
  constexpr int kBM = kTile, kD = kHeadDim;
    char* orow = p.out + b * p.o_b + (int64_t)qi * p.o_r + (int64_t)h * p.o_h;
    if (seg.parts == 1) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
      for (int c = 0; c < kD; c += 32) {
        uint32_t r[32];
        if (my_n > 0) {
          tmem_ld_x32(o_addr + c, r);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; e++) r[e] = 0;
        }
        if (i < rows_t) {
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
      if (p.lse && i < rows_t)
        p.lse[((int64_t)b * p.num_heads + h) * p.seqlen_q + qi] =
            l > 0.f ? (m_ref + log2f(l)) * 0.6931471805599453f : INFINITY;
    } else {
      // ---- split item: publish this part, the last part to arrive reduces all of them ----
      const int64_t mine = ((int64_t)seg.my_slot * 2 + t) * kBM + i;
      seg.ws_ml[mine * 2] = m_ref;
      seg.ws_ml[mine * 2 + 1] = l;
      if (my_n > 0) {
        float* dst = seg.ws_o + mine * kD;
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
          uint32_t r[32];
          tmem_ld_x32(o_addr + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<uint4*>(dst + c + e) = make_uint4(r[e], r[e + 1], r[e + 2], r[e + 3]);
        }
      }
      named_bar_sync(2, 256);  // both softmax warpgroups have stored their rows
      if (threadIdx.x == 128) {
        __threadfence();       // cumulative: orders the other threads' stores before the ticket
        bar.ticket = atomicAdd(seg.arrive, 1);
      }
      named_bar_sync(2, 256);
      if (bar.ticket == seg.parts - 1) {
        __threadfence();
        float M = -INFINITY;
        for (int c = 0; c < seg.parts; c++) {
          const int64_t sl = seg.slot_base + c;
          M = fmaxf(M, __ldcg(seg.ws_ml + ((sl * 2 + t) * kBM + i) * 2));
        }
        const float Ms = (M == -INFINITY) ? 0.f : M;
        float L = 0.f;
        for (int c = 0; c < seg.parts; c++) {
          const int64_t sl = seg.slot_base + c;
          const int64_t row = (sl * 2 + t) * kBM + i;
          const float lc = __ldcg(seg.ws_ml + row * 2 + 1);
          if (lc > 0.f) L = fmaf(lc, fast_exp2(__ldcg(seg.ws_ml + row * 2) - Ms), L);
        }
        const float inv = L > 0.f ? 1.f / L : 0.f;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {  // 64 columns at a time: 64 accumulator registers
          float acc[kD / 2];
#pragma unroll
          for (int e = 0; e < kD / 2; e++) acc[e] = 0.f;
          for (int c = 0; c < seg.parts; c++) {
            const int64_t sl = seg.slot_base + c;
            const int64_t row = (sl * 2 + t) * kBM + i;
            if (!(__ldcg(seg.ws_ml + row * 2 + 1) > 0.f)) continue;  // empty part: its O row was never written
            const float w = fast_exp2(__ldcg(seg.ws_ml + row * 2) - Ms);
            const float4* src = reinterpret_cast<const float4*>(seg.ws_o + row * kD + half * (kD / 2));
#pragma unroll
            for (int e = 0; e < kD / 8; e++) {
              const float4 v = __ldcg(src + e);
              acc[4 * e] = fmaf(v.x, w, acc[4 * e]), acc[4 * e + 1] = fmaf(v.y, w, acc[4 * e + 1]);
              acc[4 * e + 2] = fmaf(v.z, w, acc[4 * e + 2]), acc[4 * e + 3] = fmaf(v.w, w, acc[4 * e + 3]);
            }
          }
          if (i < rows_t) {
#pragma unroll
            for (int e = 0; e < kD / 2; e += 8) {
              uint4 o;
              o.x = Elem<T>::from_f2(acc[e] * inv, acc[e + 1] * inv);
              o.y = Elem<T>::from_f2(acc[e + 2] * inv, acc[e + 3] * inv);
              o.z = Elem<T>::from_f2(acc[e + 4] * inv, acc[e + 5] * inv);
              o.w = Elem<T>::from_f2(acc[e + 6] * inv, acc[e + 7] * inv);
              *reinterpret_cast<uint4*>(orow + (half * (kD / 2) + e) * 2) = o;
            }
          }
        }
        if (p.lse && i < rows_t)
          p.lse[((int64_t)b * p.num_heads + h) * p.seqlen_q + qi] =
              L > 0.f ? (M + log2f(L)) * 0.6931471805599453f : INFINITY;
        if (threadIdx.x == 128) *seg.arrive = 0;  // ready for the next launch
      }
    }
}

template <typename T, int MODE = 0, int ROLE = 0>
__device__ void prefill2_work(const CUtensorMap* qmap, const CUtensorMap* kmap, const CUtensorMap* vmap,
                              const CUtensorMap* kmap_tail, const CUtensorMap* vmap_tail,
                              const PrefillParams& p, Prefill2Smem& sm, TcBarriers& bar, uint32_t tmem,
                              const PrefillSegment& seg, bool barriers_live) {
  constexpr int kStages = kPrefill2Stages;
  constexpr int kBM = kTile, kBN = kTile, kD = kHeadDim;
  constexpr bool REGS = MODE > 0;
  const int h = seg.h, b = seg.b;
  const int hkv = h / p.group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = p.cache_batch_idx ? p.cache_batch_idx[b] : b;
  const int lk = (p.cache_seqlens ? p.cache_seqlens[b] : p.seqlen_k) + p.seqlen_new;
  const int shift = lk - p.seqlen_q;
  int m0[2], rows[2], nt[2], nl[2];
  prefill_pair_tiles(p, lk, seg.mt2, m0, rows, nt);
  const int j0 = seg.j0;
#pragma unroll
  for (int t = 0; t < 2; t++) nl[t] = max(0, min(seg.j1, nt[t]) - j0);  // local tile count of block t
  const int n = max(nl[0], nl[1]);                                      // key tiles to stage

  if (threadIdx.x == 0) {
    mbar_reinit(&bar.q_full, 1, barriers_live);
    for (int s = 0; s < kStages; s++) {
      mbar_reinit(&bar.full[s], 1, barriers_live);
      mbar_reinit(&bar.empty[s], 1, barriers_live);
    }
    for (int i = 0; i < 2; i++) {
      mbar_reinit(&bar.s_full[i], 1, barriers_live);
      mbar_reinit(&bar.p_ready[i], 128, barriers_live);
      mbar_reinit(&bar.o_full[i], 1, barriers_live);
    }
    fence_mbar_init();
  }
  cta_sync_384();

  if constexpr (ROLE != 2) {
  if (ROLE == 1 || warp < 4) {
  if (warp == 0) {
    // =========================================================== TMA producer ====
    if (lane == 0 && n > 0) {
      mbar_expect_tx(&bar.q_full, 2 * kBM * kD * 2);
      tma_load_5d(sm.q[0], qmap, &bar.q_full, 0, m0[0], 0, h, b);
      tma_load_5d(sm.q[1], qmap, &bar.q_full, 0, m0[1], 0, h, b);  // rows past seqlen_q arrive as zeros
      // consumption order K0 V0 K1 V1 ...: position 2j is K_j, 2j+1 is V_j
      const int safe_rows = (lk + p.tail_rows - 1) / p.tail_rows * p.tail_rows;
      for (int pos = 0; pos < 2 * n; pos++) {
        const int s = pos % kStages;
        mbar_wait(&bar.empty[s], ((pos / kStages) & 1) ^ 1);
        load_kv_tile(sm.ring[s], (pos & 1) ? vmap : kmap, (pos & 1) ? vmap_tail : kmap_tail, &bar.full[s],
                     (j0 + (pos >> 1)) * kBN, hkv, slot, safe_rows, p.tail_rows);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer ====
    if (lane == 0 && n > 0) {
      mbar_wait(&bar.q_full, 0);
      auto wait_slot = [&](int pos) {
        mbar_wait(&bar.full[pos % kStages], (pos / kStages) & 1);
        tc_fence_after();
        return smem_u32(sm.ring[pos % kStages]);
      };
      auto issue_qk = [&](int t, uint32_t k0) {
        const uint32_t q_addr = smem_u32(sm.q[t]);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ks++) {
          const uint32_t off = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
          umma_ss(tmem + kCol2S + t * kBN, make_smem_desc(q_addr + off, 16, 1024, kLayoutSw128),
                  make_smem_desc(k0 + off, 16, 1024, kLayoutSw128), p.idesc_qk, ks > 0);
        }
        umma_commit(&bar.s_full[t]);
      };
      auto issue_pv = [&](int t, int j, uint32_t v0) {
        mbar_wait(&bar.p_ready[t], j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < kBN / 16; ks++)
          umma_ts(tmem + kCol2O + t * kD, tmem + kCol2S + t * kBN + ks * 8,
                  make_smem_desc(v0 + ks * (16 * 128), p.v_lbo, p.v_sbo, kLayoutSw128), p.idesc_pv,
                  (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&bar.o_full[t]);  // completes phase j of block t
      };
      // prologue: S_t(0) for both blocks
      {
        const uint32_t k0 = wait_slot(0);
        if (nl[0] > 0) issue_qk(0, k0);
        if (nl[1] > 0) issue_qk(1, k0);
        umma_commit(&bar.empty[0]);
      }
      for (int j = 0; j < n; j++) {
        const uint32_t v0 = wait_slot(2 * j + 1);
        uint32_t k1 = 0;
        const bool more = j + 1 < n;
        // block 0: PV(j) then the next QK^T, which overwrites S0/P0 only after PV0(j) in order
        if (j < nl[0]) issue_pv(0, j, v0);
        if (more) {
          k1 = wait_slot(2 * j + 2);
          if (j + 1 < nl[0]) issue_qk(0, k1);
        }
        if (j < nl[1]) issue_pv(1, j, v0);
        umma_commit(&bar.empty[(2 * j + 1) % kStages]);  // V_j consumed by both blocks
        if (more) {
          if (j + 1 < nl[1]) issue_qk(1, k1);
          umma_commit(&bar.empty[(2 * j + 2) % kStages]);  // K_{j+1} consumed by both blocks
        }
      }
      // the last commit (V_{n-1}'s slot) is observed by nobody else: wait for it here, so that no
      // asynchronous arrival is still in flight when a persistent caller re-initialises the barriers
      mbar_wait(&bar.empty[(2 * n - 1) % kStages], ((2 * n - 1) / kStages) & 1);
    }
  }
  }
  }
  if constexpr (ROLE != 1) {
  if (ROLE == 2 || warp >= 4) {
    // ==================================================== softmax / epilogue ====
    const int t = (warp - 4) >> 2;            // which row block this warpgroup owns
    const int i = (threadIdx.x - 128) & 127;  // query row inside the block == TMEM lane
    const int sw = warp & 3;
    const uint32_t lane_base = (uint32_t)(sw * 32) << 16;
    const int qi = m0[t] + i;
    const int my_n = nl[t];
    int limit = lk - 1;
    if (p.causal) limit = min(limit, qi + shift);
    float m_ref = -INFINITY, l = 0.f;
    const uint32_t s_addr = tmem + lane_base + kCol2S + t * kBN;
    const uint32_t o_addr = tmem + lane_base + kCol2O + t * kD;
    for (int j = 0; j < my_n; j++) {
      mbar_wait(&bar.s_full[t], j & 1);
      tc_fence_after();
      const int key0 = (j0 + j) * kBN;
      const bool need_mask = key0 + kBN - 1 > limit;
      const bool warp_mask = __any_sync(0xffffffffu, need_mask);
      uint32_t s0[REGS ? 32 : 1], s1[REGS ? 32 : 1], s2[REGS ? 32 : 1], s3[REGS ? 32 : 1];
      float mx;
      if constexpr (REGS) {
        tmem_ld_x32(s_addr, s0);
        tmem_ld_x32(s_addr + 32, s1);
        tmem_ld_x32(s_addr + 64, s2);
        tmem_ld_x32(s_addr + 96, s3);
        tmem_wait_ld();
        if (warp_mask)
          mx = fmaxf(fmaxf(regs_quarter_max<true>(s0, key0, limit), regs_quarter_max<true>(s1, key0 + 32, limit)),
                     fmaxf(regs_quarter_max<true>(s2, key0 + 64, limit), regs_quarter_max<true>(s3, key0 + 96, limit)));
        else
          mx = fmaxf(fmaxf(regs_quarter_max<false>(s0, 0, 0), regs_quarter_max<false>(s1, 0, 0)),
                     fmaxf(regs_quarter_max<false>(s2, 0, 0), regs_quarter_max<false>(s3, 0, 0)));
      } else {
        mx = warp_mask ? tile_row_max<true>(s_addr, key0, limit) : tile_row_max<false>(s_addr, key0, limit);
      }
      mx *= p.scale_log2;
      float alpha = 1.f;
      bool grow = mx > m_ref + kRescaleThreshold;
      if (m_ref == -INFINITY && mx > -INFINITY) grow = true;
      if (grow) {
        alpha = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - mx);
        m_ref = mx;
      }
      if (__any_sync(0xffffffffu, grow) && j > 0) {
        // S_t(j) is ready, so PV_t(j-1) -- issued before this tile's QK^T -- has retired
        mbar_wait(&bar.o_full[t], (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
          uint32_t r[32];
          tmem_ld_x32(o_addr + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
          tmem_st_x32(o_addr + c, r);
        }
        tmem_wait_st();
      }
      l *= alpha;
      const float mref_safe = (m_ref == -INFINITY) ? 0.f : m_ref;
      // P_t(j) overwrites the already consumed low half of S_t (in place)
      if constexpr (REGS) {
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t packed[32];
        if (warp_mask) {
          regs_quarter_exp<T, true>(s0, packed, p.scale_log2, mref_safe, key0, limit, ls);
          regs_quarter_exp<T, true>(s1, packed + 16, p.scale_log2, mref_safe, key0 + 32, limit, ls);
        } else {
          regs_quarter_exp<T, false>(s0, packed, p.scale_log2, mref_safe, 0, 0, ls);
          regs_quarter_exp<T, false>(s1, packed + 16, p.scale_log2, mref_safe, 0, 0, ls);
        }
        tmem_st_x32(s_addr, packed);  // keys 0..63 of P_t(j), 2 per column
        if (warp_mask) {
          regs_quarter_exp<T, true>(s2, packed, p.scale_log2, mref_safe, key0 + 64, limit, ls);
          regs_quarter_exp<T, true>(s3, packed + 16, p.scale_log2, mref_safe, key0 + 96, limit, ls);
        } else {
          regs_quarter_exp<T, false>(s2, packed, p.scale_log2, mref_safe, 0, 0, ls);
          regs_quarter_exp<T, false>(s3, packed + 16, p.scale_log2, mref_safe, 0, 0, ls);
        }
        tmem_st_x32(s_addr + 32, packed);  // keys 64..127
        l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      } else {
        l += warp_mask ? tile_exp_store<T, true>(s_addr, s_addr, p.scale_log2, mref_safe, key0, limit)
                       : tile_exp_store<T, false>(s_addr, s_addr, p.scale_log2, mref_safe, key0, limit);
      }
      tmem_wait_st();
      if ((j0 + j + 1) * kBN > lk) {
        const int pv = 2 * j + 1;
        mbar_wait(&bar.full[pv % kStages], (pv / kStages) & 1);
        // both warpgroups may reach the tail tile: zeroing the same rows twice is harmless
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
      mbar_arrive(&bar.p_ready[t]);
    }

    if (my_n > 0) {
      mbar_wait(&bar.o_full[t], (my_n - 1) & 1);  // phases <= my_n-2 are known complete (see above)
      tc_fence_after();
    }
    prefill_block_epilogue<T>(p, seg, bar, tmem, o_addr, t, i, qi, h, b, rows[t], my_n, m_ref, l);
  }
  }
  tc_fence_before();
  cta_sync_384();
}

// (A variant with 64-key tiles and TWO S buffers per row block -- QK^T of tile j+2 issued right behind
// PV of tile j, so that S(j+1) is already there when the softmax of tile j ends -- was built, parity
// tested and measured in round 2: 765 / 635 / 806 TFLOP/s at chunk 2048 / 512 / 8192 against 975 / 788 /
// 1057 for prefill2; ncu: tensor pipe 37 %, XU 37 %, softmax warps waiting for S 40 % of their samples
// (profiles/r2_prefill3_discarded_ncu_*).  One tile of look-ahead at half the tile size does not cover
// the ~500-cycle round trip P -> MMA thread -> tensor pipe -> commit -> softmax; removed.)

}  // namespace tcwork
}  // namespace vattn
