// cache_flat and the KV-append used by flash_attn_with_kvcache(k=, v=).
//
// Reference: sarathi-lean/csrc/cache_kernels.cu:482-570 (cache_flat: one block per
// token, scalar 2-byte __ldg loads) and the append-KV prologue of FA's split
// kernel, pod_attn/pod_attn/flash_fwd_kernel.h:685-790 (rotary unused by the
// sarathi wrappers).  Both are pure HBM-bound copies: algorithmic bytes =
// 4 * itemsize * c * Hkv * D for cache_flat.  Here every thread moves
// 16 bytes per access, K and V in the same pass, grid sized in SM multiples.
#include "attn_common.cuh"

namespace vattn {

namespace {

constexpr int kThreads = 256;

// rows of `chunks_per_row` 16-byte chunks; strides in bytes
__global__ void __launch_bounds__(kThreads)
cache_flat_vec_kernel(const char* __restrict__ key, const char* __restrict__ value,
                      char* __restrict__ k_cache, char* __restrict__ v_cache, int64_t num_tokens,
                      int chunks_per_row, int64_t key_stride, int64_t value_stride,
                      int64_t kc_stride, int64_t vc_stride) {
  const int64_t total = num_tokens * chunks_per_row;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * kThreads) {
    const int64_t t = i / chunks_per_row;
    const int64_t c = (i - t * chunks_per_row) * 16;
    uint4 kk = ld_stream_128(key + t * key_stride + c);
    uint4 vv = ld_stream_128(value + t * value_stride + c);
    st_stream_128(k_cache + t * kc_stride + c, kk);
    st_stream_128(v_cache + t * vc_stride + c, vv);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cache_flat_scalar_kernel(const T* __restrict__ key, const T* __restrict__ value,
                         T* __restrict__ k_cache, T* __restrict__ v_cache, int64_t num_tokens,
                         int64_t row_elems, int64_t key_stride, int64_t value_stride,
                         int64_t kc_stride, int64_t vc_stride) {
  const int64_t total = num_tokens * row_elems;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * kThreads) {
    const int64_t t = i / row_elems;
    const int64_t c = i - t * row_elems;
    k_cache[t * kc_stride + c] = key[t * key_stride + c];
    v_cache[t * vc_stride + c] = value[t * value_stride + c];
  }
}

// k_new/v_new [batch, seqlen_new, Hkv, D] -> cache rows [L0, L0+seqlen_new) of slot
// (flash_fwd_kernel.h:685-790 semantics).  One 16-byte chunk per thread.
__global__ void __launch_bounds__(kThreads)
append_kv_kernel(const char* __restrict__ k_new, const char* __restrict__ v_new,
                 char* __restrict__ k_cache, char* __restrict__ v_cache,
                 const int32_t* __restrict__ cache_seqlens,
                 const int32_t* __restrict__ cache_batch_idx, int batch, int seqlen_new,
                 int num_kv_heads, int chunks_per_head, int seqlen_k,
                 int64_t kn_b, int64_t kn_r, int64_t kn_h, int64_t vn_b, int64_t vn_r, int64_t vn_h,
                 int64_t kc_b, int64_t kc_r, int64_t kc_h, int64_t vc_b, int64_t vc_r,
                 int64_t vc_h) {
  const int64_t per_b = (int64_t)seqlen_new * num_kv_heads * chunks_per_head;
  const int64_t total = per_b * batch;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * kThreads) {
    const int b = (int)(i / per_b);
    int64_t r = i - (int64_t)b * per_b;
    const int t = (int)(r / ((int64_t)num_kv_heads * chunks_per_head));
    r -= (int64_t)t * num_kv_heads * chunks_per_head;
    const int h = (int)(r / chunks_per_head);
    const int64_t c = (r - (int64_t)h * chunks_per_head) * 16;
    const int slot = cache_batch_idx ? cache_batch_idx[b] : b;
    const int row = (cache_seqlens ? cache_seqlens[b] : seqlen_k) + t;
    uint4 kk = ld_stream_128(k_new + b * kn_b + t * kn_r + h * kn_h + c);
    uint4 vv = ld_stream_128(v_new + b * vn_b + t * vn_r + h * vn_h + c);
    *reinterpret_cast<uint4*>(k_cache + slot * kc_b + row * kc_r + h * kc_h + c) = kk;
    *reinterpret_cast<uint4*>(v_cache + slot * vc_b + row * vc_r + h * vc_h + c) = vv;
  }
}

inline int grid_for(int64_t total_threads) {
  int64_t blocks = (total_threads + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)num_sms() * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

void launch_cache_flat(const void* key, const void* value, void* k_cache, void* v_cache,
                       int64_t num_tokens, int64_t row_elems, int64_t key_stride,
                       int64_t value_stride, int64_t kc_stride, int64_t vc_stride, int elem_bytes,
                       cudaStream_t stream) {
  if (num_tokens <= 0 || row_elems <= 0) return;
  if (elem_bytes != 2 && elem_bytes != 4) throw ArgError("[vattn] cache_flat: fp16/bf16/fp32 only");
  const int64_t row_bytes = row_elems * elem_bytes;
  const bool vec = row_bytes % 16 == 0 && aligned16(key) && aligned16(value) && aligned16(k_cache) &&
                   aligned16(v_cache) && (key_stride * elem_bytes) % 16 == 0 &&
                   (value_stride * elem_bytes) % 16 == 0 && (kc_stride * elem_bytes) % 16 == 0 &&
                   (vc_stride * elem_bytes) % 16 == 0;
  if (vec) {
    const int cpr = (int)(row_bytes / 16);
    cache_flat_vec_kernel<<<grid_for(num_tokens * cpr), kThreads, 0, stream>>>(
        (const char*)key, (const char*)value, (char*)k_cache, (char*)v_cache, num_tokens, cpr,
        key_stride * elem_bytes, value_stride * elem_bytes, kc_stride * elem_bytes,
        vc_stride * elem_bytes);
  } else if (elem_bytes == 2) {
    cache_flat_scalar_kernel<uint16_t><<<grid_for(num_tokens * row_elems), kThreads, 0, stream>>>(
        (const uint16_t*)key, (const uint16_t*)value, (uint16_t*)k_cache, (uint16_t*)v_cache,
        num_tokens, row_elems, key_stride, value_stride, kc_stride, vc_stride);
  } else {
    cache_flat_scalar_kernel<uint32_t><<<grid_for(num_tokens * row_elems), kThreads, 0, stream>>>(
        (const uint32_t*)key, (const uint32_t*)value, (uint32_t*)k_cache, (uint32_t*)v_cache,
        num_tokens, row_elems, key_stride, value_stride, kc_stride, vc_stride);
  }
  count_launch();
  VATTN_CUDA(cudaGetLastError());
}

void launch_append_kv(const vattn_fwd_params_t& p, cudaStream_t stream) {
  if (!p.k_new || p.seqlen_new <= 0 || p.batch <= 0) return;
  const int eb = 2;
  if ((p.head_dim * eb) % 16 != 0) throw ArgError("[vattn] head_dim must be a multiple of 8");
  const int cph = p.head_dim * eb / 16;
  const int64_t total = (int64_t)p.batch * p.seqlen_new * p.num_kv_heads * cph;
  append_kv_kernel<<<grid_for(total), kThreads, 0, stream>>>(
      (const char*)p.k_new, (const char*)p.v_new, (char*)p.k_cache, (char*)p.v_cache,
      p.cache_seqlens, p.cache_batch_idx, p.batch, p.seqlen_new, p.num_kv_heads, cph, p.seqlen_k,
      p.knew_batch_stride * eb, p.knew_row_stride * eb, p.knew_head_stride * eb,
      p.vnew_batch_stride * eb, p.vnew_row_stride * eb, p.vnew_head_stride * eb,
      p.k_batch_stride * eb, p.k_row_stride * eb, p.k_head_stride * eb, p.v_batch_stride * eb,
      p.v_row_stride * eb, p.v_head_stride * eb);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
}

}  // namespace vattn
