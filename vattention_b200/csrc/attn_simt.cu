// SIMT attention over contiguous K/V: a coalesced 128-bit-vectorised HBM sweep with
// warp-shuffle softmax reductions.  This is the generic path (any seqlen_q,
// head_dim 64/128, any GQA ratio, fp16/bf16) and the parity anchor for the
// tcgen05 kernels; the tensor-core paths take over where the shapes allow
// (see capi_attn.cu for why CUDA-core FMAs cannot sustain GQA decode at HBM
// speed on B200).
//
// Semantics restated from the FA-2 fork the reference vendors:
//   key range / append     pod_attn/pod_attn/block_info.h:11-44
//   causal limit           pod_attn/pod_attn/mask.h:172  (col < row + 1 + Lk - Sq)
//   exp2 softmax, -inf     pod_attn/pod_attn/softmax.h:66-160
//   split-KV + combine     pod_attn/pod_attn/flash_fwd_kernel.h:503-1077, 1115+
//   GQA head mapping       h_kv = h_q / (Hq/Hkv)   (flash_api.cpp:1370-1378)
//
// Work decomposition: one CTA of 4 warps per (split, kv-head chunk, q row).  All
// q heads of the GQA group are processed against each K/V byte, so K/V is read
// once per group.  Each half-warp (D=128) owns whole rows: a 16-byte load per
// lane covers 8 dims; the dot product is reduced over the 16 lanes of the row.
#include "attn_common.cuh"

namespace vattn {

namespace {

constexpr int kWarps = 4;
constexpr int kUnroll = 4;

struct SimtParams {
  const char* q;
  const char* k;
  const char* v;
  char* out;
  float* lse;
  float* ws_acc;
  float* ws_ml;
  const int32_t* cache_seqlens;
  const int32_t* cache_batch_idx;
  int64_t q_b, q_r, q_h;  // byte strides
  int64_t k_b, k_r, k_h;
  int64_t v_b, v_r, v_h;
  int64_t o_b, o_r, o_h;
  int batch, seqlen_q, seqlen_k, seqlen_new, num_heads, num_kv_heads, group;
  int chunks_per_group;  // ceil(group / GQ)
  int causal;
  int num_splits;
  float scale_log2;
};

template <typename T, int D, int GQ>
__global__ void __launch_bounds__(kWarps * 32)
attn_simt_kernel(const SimtParams p) {
  constexpr int LPR = D / 8;        // lanes per row (16-byte chunks in a row)
  constexpr int RPL = 32 / LPR;     // rows per warp-wide load
  constexpr int RPI = RPL * kUnroll;  // rows per warp per iteration

  // query rows on grid.x (up to 2^31 - 1 blocks: a 128K-token prefill has more rows than the 65535
  // that grid.y / grid.z allow), splits on grid.z (<= 128)
  const int split = blockIdx.z;
  const int hkv = blockIdx.y / p.chunks_per_group;
  const int hchunk = blockIdx.y % p.chunks_per_group;
  const int row_id = blockIdx.x;  // b * seqlen_q + i
  const int b = row_id / p.seqlen_q;
  const int qi = row_id - b * p.seqlen_q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR;  // which row of the warp-wide load this lane reads
  const int dl = lane % LPR;   // which 16-byte chunk of the row

  const int h0 = hkv * p.group + hchunk * GQ;  // first q head of this chunk
  const int nvalid = min(GQ, p.group - hchunk * GQ);

  const int slot = p.cache_batch_idx ? p.cache_batch_idx[b] : b;
  const int lk = (p.cache_seqlens ? p.cache_seqlens[b] : p.seqlen_k) + p.seqlen_new;
  int visible = lk;
  if (p.causal) visible = min(lk, qi + lk - p.seqlen_q + 1);  // mask.h:172
  if (visible < 0) visible = 0;
  // split boundaries depend on lk only, so every q row of a batch entry agrees
  int per_split = (lk + p.num_splits - 1) / p.num_splits;
  per_split = (per_split + 63) / 64 * 64;
  const int k_begin = split * per_split;
  const int k_end = min(k_begin + per_split, visible);

  // ---- load q (pre-scaled into the log2 domain) ----
  float q[GQ][8];
#pragma unroll
  for (int g = 0; g < GQ; g++) {
    if (g < nvalid) {
      const uint4 u = *reinterpret_cast<const uint4*>(p.q + b * p.q_b + qi * p.q_r +
                                                      (int64_t)(h0 + g) * p.q_h + dl * 16);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float2 f = Elem<T>::to_f2(w[j]);
        q[g][2 * j] = f.x * p.scale_log2;
        q[g][2 * j + 1] = f.y * p.scale_log2;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) q[g][j] = 0.f;
    }
  }

  float o[GQ][8], m[GQ], l[GQ];
#pragma unroll
  for (int g = 0; g < GQ; g++) {
    m[g] = -INFINITY;
    l[g] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) o[g][j] = 0.f;
  }

  const char* kbase = p.k + (int64_t)slot * p.k_b + (int64_t)hkv * p.k_h + dl * 16;
  const char* vbase = p.v + (int64_t)slot * p.v_b + (int64_t)hkv * p.v_h + dl * 16;

  for (int r0 = k_begin + warp * RPI; r0 < k_end; r0 += kWarps * RPI) {
    uint4 kk[kUnroll], vv[kUnroll];
    bool ok[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      const int row = r0 + u * RPL + sub;
      ok[u] = row < k_end;
      // rows past the mapped prefix are unmapped VA: never form the access
      kk[u] = ok[u] ? ld_stream_128(kbase + (int64_t)row * p.k_r) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      const int row = r0 + u * RPL + sub;
      vv[u] = ok[u] ? ld_stream_128(vbase + (int64_t)row * p.v_r) : make_uint4(0, 0, 0, 0);
    }

    float s[kUnroll][GQ];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      const uint32_t w[4] = {kk[u].x, kk[u].y, kk[u].z, kk[u].w};
      float kf[8];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float2 f = Elem<T>::to_f2(w[j]);
        kf[2 * j] = f.x;
        kf[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int g = 0; g < GQ; g++) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) acc = fmaf(q[g][j], kf[j], acc);
        s[u][g] = acc;
      }
    }
    // reduce each dot product over the LPR lanes that share a row
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) {
#pragma unroll
      for (int u = 0; u < kUnroll; u++)
#pragma unroll
        for (int g = 0; g < GQ; g++) s[u][g] += __shfl_xor_sync(0xffffffffu, s[u][g], off);
    }

#pragma unroll
    for (int g = 0; g < GQ; g++) {
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
        if (!ok[u]) s[u][g] = -INFINITY;
        mx = fmaxf(mx, s[u][g]);
      }
      const float m_new = fmaxf(m[g], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;  // softmax.h:76-78
      const float alpha = fast_exp2(m[g] - m_safe);
      m[g] = m_new;
      float psum = 0.f;
      float pv[8];
#pragma unroll
      for (int j = 0; j < 8; j++) pv[j] = 0.f;
#pragma unroll
      for (int u = 0; u < kUnroll; u++) {
        const float pr = fast_exp2(s[u][g] - m_safe);
        psum += pr;
        const uint32_t w[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float2 f = Elem<T>::to_f2(w[j]);
          pv[2 * j] = fmaf(pr, f.x, pv[2 * j]);
          pv[2 * j + 1] = fmaf(pr, f.y, pv[2 * j + 1]);
        }
      }
      l[g] = l[g] * alpha + psum;
#pragma unroll
      for (int j = 0; j < 8; j++) o[g][j] = fmaf(o[g][j], alpha, pv[j]);
    }
  }

  // ---- merge the RPL independent row streams of this warp ----
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
    for (int g = 0; g < GQ; g++) {
      const float m_o = __shfl_xor_sync(0xffffffffu, m[g], off);
      const float l_o = __shfl_xor_sync(0xffffffffu, l[g], off);
      const float m_new = fmaxf(m[g], m_o);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float a = fast_exp2(m[g] - m_safe), bsc = fast_exp2(m_o - m_safe);
      l[g] = l[g] * a + l_o * bsc;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float o_o = __shfl_xor_sync(0xffffffffu, o[g][j], off);
        o[g][j] = o[g][j] * a + o_o * bsc;
      }
      m[g] = m_new;
    }
  }

  // ---- merge the warps through shared memory ----
  __shared__ float sm_o[kWarps][GQ][D];
  __shared__ float sm_ml[kWarps][GQ][2];
  if (sub == 0) {
#pragma unroll
    for (int g = 0; g < GQ; g++) {
#pragma unroll
      for (int j = 0; j < 8; j++) sm_o[warp][g][dl * 8 + j] = o[g][j];
      if (dl == 0) {
        sm_ml[warp][g][0] = m[g];
        sm_ml[warp][g][1] = l[g];
      }
    }
  }
  __syncthreads();

  const int64_t row64 = row_id;
  for (int idx = threadIdx.x; idx < GQ * D; idx += kWarps * 32) {
    const int g = idx / D, d = idx - g * D;
    if (g >= nvalid) continue;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kWarps; w++) M = fmaxf(M, sm_ml[w][g][0]);
    const float Ms = (M == -INFINITY) ? 0.f : M;
    float L = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const float sc = fast_exp2(sm_ml[w][g][0] - Ms);
      L += sm_ml[w][g][1] * sc;
      acc += sm_o[w][g][d] * sc;
    }
    const int h = h0 + g;
    if (p.num_splits == 1) {
      const float inv = (L > 0.f) ? 1.f / L : 0.f;  // no visible key -> zeros
      reinterpret_cast<T*>(p.out + b * p.o_b + qi * p.o_r + (int64_t)h * p.o_h)[d] =
          Elem<T>::from_f(acc * inv);
      if (p.lse && d == 0)
        p.lse[((int64_t)b * p.num_heads + h) * p.seqlen_q + qi] =
            (L > 0.f) ? (M + log2f(L)) * 0.6931471805599453f : INFINITY;
    } else {
      const int64_t base = ((row64 * p.num_heads + h) * p.num_splits + split);
      p.ws_acc[base * D + d] = acc;
      if (d == 0) {
        p.ws_ml[base * 2] = M;
        p.ws_ml[base * 2 + 1] = L;
      }
    }
  }
}

// LSE-weighted reduction of the split partials (flash_fwd_kernel.h:1115+).
// One warp per (row, head); each lane owns 4 consecutive dims (one float4 per split, loads of
// successive splits independent of each other).
template <typename T>
__global__ void __launch_bounds__(128)
combine_kernel(const float* __restrict__ ws_acc, const float* __restrict__ ws_ml,
               char* __restrict__ out, float* __restrict__ lse, int64_t total_rows_heads,
               int num_heads, int seqlen_q, int num_splits, int head_dim, int64_t o_b, int64_t o_r,
               int64_t o_h) {
  const int64_t rh = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (rh >= total_rows_heads) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = rh / num_heads;
  const int h = (int)(rh - row * num_heads);
  const int64_t b = row / seqlen_q;
  const int qi = (int)(row - b * seqlen_q);
  const float* ml = ws_ml + rh * num_splits * 2;
  float M = -INFINITY;
  for (int s = lane; s < num_splits; s += 32) M = fmaxf(M, ml[2 * s]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, off));
  const float Ms = (M == -INFINITY) ? 0.f : M;
  float L = 0.f;
  for (int s = lane; s < num_splits; s += 32) L += ml[2 * s + 1] * fast_exp2(ml[2 * s] - Ms);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) L += __shfl_xor_sync(0xffffffffu, L, off);
  const float inv = (L > 0.f) ? 1.f / L : 0.f;
  const float4* acc = reinterpret_cast<const float4*>(ws_acc + rh * num_splits * head_dim);
  T* o = reinterpret_cast<T*>(out + b * o_b + qi * o_r + (int64_t)h * o_h);
  const int d4n = head_dim / 4;  // float4 columns; head_dim is a multiple of 8
  for (int d4b = 0; d4b < d4n; d4b += 32) {  // warp-uniform trip count (shuffles inside)
    const int d4 = d4b + lane;
    const bool active = d4 < d4n;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < num_splits; s0 += 32) {
      // lane j holds the weight of split s0 + j; 0 for empty partials, whose acc was never
      // written and must not be read (0 * garbage)
      const int sj = s0 + lane;
      const float w_mine = (sj < num_splits && ml[2 * sj + 1] > 0.f) ? fast_exp2(ml[2 * sj] - Ms) : 0.f;
      const int cnt = min(32, num_splits - s0);
#pragma unroll 4
      for (int j = 0; j < cnt; j++) {
        const float w = __shfl_sync(0xffffffffu, w_mine, j);
        if (w != 0.f && active) {
          const float4 v = acc[(int64_t)(s0 + j) * d4n + d4];
          a.x = fmaf(v.x, w, a.x), a.y = fmaf(v.y, w, a.y), a.z = fmaf(v.z, w, a.z), a.w = fmaf(v.w, w, a.w);
        }
      }
    }
    uint2 packed;
    packed.x = Elem<T>::from_f2(a.x * inv, a.y * inv);
    packed.y = Elem<T>::from_f2(a.z * inv, a.w * inv);
    if (active) *reinterpret_cast<uint2*>(o + d4 * 4) = packed;
  }
  if (lse && lane == 0)
    lse[(b * num_heads + h) * seqlen_q + qi] =
        (L > 0.f) ? (M + log2f(L)) * 0.6931471805599453f : INFINITY;
}

template <typename T, int D>
void launch_simt_t(const vattn_fwd_params_t& p, int splits, const SplitWorkspace& ws,
                   cudaStream_t stream) {
  SimtParams sp;
  const int eb = 2;
  sp.q = (const char*)p.q;
  sp.k = (const char*)p.k_cache;
  sp.v = (const char*)p.v_cache;
  sp.out = (char*)p.out;
  sp.lse = p.softmax_lse;
  sp.ws_acc = ws.acc;
  sp.ws_ml = ws.ml;
  sp.cache_seqlens = p.cache_seqlens;
  sp.cache_batch_idx = p.cache_batch_idx;
  sp.q_b = p.q_batch_stride * eb, sp.q_r = p.q_row_stride * eb, sp.q_h = p.q_head_stride * eb;
  sp.k_b = p.k_batch_stride * eb, sp.k_r = p.k_row_stride * eb, sp.k_h = p.k_head_stride * eb;
  sp.v_b = p.v_batch_stride * eb, sp.v_r = p.v_row_stride * eb, sp.v_h = p.v_head_stride * eb;
  sp.o_b = p.o_batch_stride * eb, sp.o_r = p.o_row_stride * eb, sp.o_h = p.o_head_stride * eb;
  sp.batch = p.batch, sp.seqlen_q = p.seqlen_q, sp.seqlen_k = p.seqlen_k;
  sp.seqlen_new = p.k_new ? p.seqlen_new : 0;
  sp.num_heads = p.num_heads, sp.num_kv_heads = p.num_kv_heads;
  sp.group = p.num_heads / p.num_kv_heads;
  sp.causal = p.causal;
  sp.num_splits = splits;
  sp.scale_log2 = p.softmax_scale * kLog2e;
  const int gq = sp.group >= 8 ? 8 : (sp.group > 2 ? 4 : sp.group);
  sp.chunks_per_group = (sp.group + gq - 1) / gq;
  dim3 grid(p.batch * p.seqlen_q, p.num_kv_heads * sp.chunks_per_group, splits);
  dim3 block(kWarps * 32);
  const int tslot = timing_begin(stream);
  switch (gq) {
    case 1: attn_simt_kernel<T, D, 1><<<grid, block, 0, stream>>>(sp); break;
    case 2: attn_simt_kernel<T, D, 2><<<grid, block, 0, stream>>>(sp); break;
    case 4: attn_simt_kernel<T, D, 4><<<grid, block, 0, stream>>>(sp); break;
    default: attn_simt_kernel<T, D, 8><<<grid, block, 0, stream>>>(sp); break;
  }
  timing_end(tslot, stream);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
}

}  // namespace

void launch_combine(const vattn_fwd_params_t& p, int splits, const SplitWorkspace& ws,
                    cudaStream_t stream) {
  const int64_t rh = (int64_t)p.batch * p.seqlen_q * p.num_heads;
  const int blocks = (int)((rh + 3) / 4);
  if (p.dtype == VATTN_DTYPE_BF16)
    combine_kernel<__nv_bfloat16><<<blocks, 128, 0, stream>>>(
        ws.acc, ws.ml, (char*)p.out, p.softmax_lse, rh, p.num_heads, p.seqlen_q, splits, p.head_dim,
        p.o_batch_stride * 2, p.o_row_stride * 2, p.o_head_stride * 2);
  else
    combine_kernel<__half><<<blocks, 128, 0, stream>>>(
        ws.acc, ws.ml, (char*)p.out, p.softmax_lse, rh, p.num_heads, p.seqlen_q, splits, p.head_dim,
        p.o_batch_stride * 2, p.o_row_stride * 2, p.o_head_stride * 2);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
}

// number of KV splits for the SIMT path: enough CTAs for ~4 waves of 148 SMs x 4
// resident CTAs, never splitting below 256 keys (cf. num_splits_heuristic,
// pod_attn/pod_attn/flash_api.cpp:258-292, which targets SM occupancy too)
int simt_num_splits(const vattn_fwd_params_t& p) {
  if (p.num_splits > 0) return p.num_splits;
  const int group = p.num_heads / p.num_kv_heads;
  const int gq = group >= 8 ? 8 : (group > 2 ? 4 : group);
  const int64_t base = (int64_t)p.batch * p.seqlen_q * p.num_kv_heads * ((group + gq - 1) / gq);
  const int64_t target = (int64_t)num_sms() * 4 * 4;
  int64_t s = (target + base - 1) / base;
  const int max_by_len = (p.seqlen_k + (p.k_new ? p.seqlen_new : 0) + 255) / 256;
  if (s > max_by_len) s = max_by_len;
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return (int)s;
}

void launch_simt(const vattn_fwd_params_t& p, int splits, const SplitWorkspace& ws,
                 cudaStream_t stream) {
  if (p.head_dim != 64 && p.head_dim != 128)
    throw UnsupportedError("[vattn] head_dim must be 64 or 128");
  if (p.dtype == VATTN_DTYPE_BF16) {
    if (p.head_dim == 128) launch_simt_t<__nv_bfloat16, 128>(p, splits, ws, stream);
    else launch_simt_t<__nv_bfloat16, 64>(p, splits, ws, stream);
  } else {
    if (p.head_dim == 128) launch_simt_t<__half, 128>(p, splits, ws, stream);
    else launch_simt_t<__half, 64>(p, splits, ws, stream);
  }
  if (splits > 1) launch_combine(p, splits, ws, stream);
}

}  // namespace vattn
