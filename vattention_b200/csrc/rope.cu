// Rotary embedding of q and of the appended keys, as the prologue of
// flash_attn_with_kvcache(k=, v=, rotary_cos=, rotary_sin=).
//
// Reference semantics: pod_attn/pod_attn/flash_api.cpp:1503-1527 (argument rules) and
// flash_fwd_kernel.h:684-830 (where the rotation happens):
//   * new key row t of batch entry b sits at position cache_seqlens[b] + t;
//   * query row i sits at cache_seqlens[b] + i when causal, at cache_seqlens[b] otherwise
//     (":796-804": the cos/sin row stride is 0 for the non-causal case);
//   * interleaved (GPT-J) pairs dims (2j, 2j+1), contiguous (NeoX) pairs (j, j + rotary_dim/2);
//     dims >= rotary_dim pass through;
//   * x0' = x0 cos - x1 sin, x1' = x0 sin + x1 cos in fp32, rounded once to the storage type.
// The rotated copies go to the call's workspace; the attention kernels then read q from there
// and append the rotated keys (the decode kernel fuses that append, attn_tc_work.cuh).  Traffic
// is 2 * itemsize * B * (Sq*Hq + Snew*Hkv) * D bytes, three orders of magnitude below the K/V
// sweep: HBM/launch-latency bound, one 16-byte vector per thread.
#include "attn_common.cuh"

namespace vattn {

namespace {

constexpr int kThreads = 256;

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = Elem<T>::to_f2(u.x), b = Elem<T>::to_f2(u.y), c = Elem<T>::to_f2(u.z), d = Elem<T>::to_f2(u.w);
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = Elem<T>::from_f2(f[0], f[1]);
  u.y = Elem<T>::from_f2(f[2], f[3]);
  u.z = Elem<T>::from_f2(f[4], f[5]);
  u.w = Elem<T>::from_f2(f[6], f[7]);
  return u;
}

// one 16-byte vector (8 elements at dims [d0, d0+8)) of one (row, head) of src -> dst
template <typename T>
__device__ __forceinline__ void rotate_vec(const T* __restrict__ src, T* __restrict__ dst, int d0,
                                           const T* __restrict__ cos_row, const T* __restrict__ sin_row,
                                           int rotary_dim, bool interleaved) {
  uint4 x = *reinterpret_cast<const uint4*>(src + d0);
  if (d0 >= rotary_dim) {
    *reinterpret_cast<uint4*>(dst + d0) = x;
    return;
  }
  float xf[8], of[8];
  unpack8<T>(x, xf);
  if (interleaved) {
    // 4 pairs; cos/sin entries d0/2 .. d0/2+3 (8 bytes each)
    const uint2 cu = *reinterpret_cast<const uint2*>(cos_row + d0 / 2);
    const uint2 su = *reinterpret_cast<const uint2*>(sin_row + d0 / 2);
    const float2 c01 = Elem<T>::to_f2(cu.x), c23 = Elem<T>::to_f2(cu.y);
    const float2 s01 = Elem<T>::to_f2(su.x), s23 = Elem<T>::to_f2(su.y);
    const float c[4] = {c01.x, c01.y, c23.x, c23.y}, s[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      of[2 * j] = xf[2 * j] * c[j] - xf[2 * j + 1] * s[j];
      of[2 * j + 1] = xf[2 * j] * s[j] + xf[2 * j + 1] * c[j];
    }
  } else {
    const int half = rotary_dim / 2;  // a multiple of 8: the partner vector is aligned too
    const bool low = d0 < half;
    const int j0 = low ? d0 : d0 - half;
    float pf[8], c[8], s[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(src + (low ? d0 + half : d0 - half)), pf);
    unpack8<T>(*reinterpret_cast<const uint4*>(cos_row + j0), c);
    unpack8<T>(*reinterpret_cast<const uint4*>(sin_row + j0), s);
#pragma unroll
    for (int j = 0; j < 8; j++)
      of[j] = low ? xf[j] * c[j] - pf[j] * s[j] : pf[j] * s[j] + xf[j] * c[j];
  }
  *reinterpret_cast<uint4*>(dst + d0) = pack8<T>(of);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
rope_qk_kernel(const T* __restrict__ q, int64_t q_b, int64_t q_r, int64_t q_h, T* __restrict__ q_out,
               const T* __restrict__ k, int64_t k_b, int64_t k_r, int64_t k_h, T* __restrict__ k_out,
               const T* __restrict__ cos_tab, const T* __restrict__ sin_tab,
               const int32_t* __restrict__ cache_seqlens, int batch, int seqlen_q, int seqlen_new,
               int num_heads, int num_kv_heads, int head_dim, int seqlen_k, int rotary_dim,
               int interleaved, int causal) {
  const int vpr = head_dim / 8;
  const int64_t q_vecs = (int64_t)batch * seqlen_q * num_heads * vpr;
  const int64_t k_vecs = (int64_t)batch * seqlen_new * num_kv_heads * vpr;
  const int half = rotary_dim / 2;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < q_vecs + k_vecs;
       i += (int64_t)gridDim.x * kThreads) {
    const bool is_q = i < q_vecs;
    int64_t r = is_q ? i : i - q_vecs;
    const int rows = is_q ? seqlen_q : seqlen_new, heads = is_q ? num_heads : num_kv_heads;
    const int d0 = (int)(r % vpr) * 8;
    r /= vpr;
    const int h = (int)(r % heads);
    r /= heads;
    const int t = (int)(r % rows);
    const int b = (int)(r / rows);
    const int base = cache_seqlens ? cache_seqlens[b] : seqlen_k;
    const int pos = base + ((is_q && !causal) ? 0 : t);
    const T* src = is_q ? q + b * q_b + t * q_r + h * q_h : k + b * k_b + t * k_r + h * k_h;
    T* dst = (is_q ? q_out : k_out) + (((int64_t)b * rows + t) * heads + h) * head_dim;
    rotate_vec<T>(src, dst, d0, cos_tab + (int64_t)pos * half, sin_tab + (int64_t)pos * half,
                  rotary_dim, interleaved != 0);
  }
}

}  // namespace

size_t rope_workspace_bytes(const vattn_fwd_params_t& p) {
  if (!p.rotary_cos || p.rotary_dim <= 0) return 0;
  const size_t q = (size_t)p.batch * p.seqlen_q * p.num_heads * p.head_dim * 2;
  const size_t k = (size_t)p.batch * p.seqlen_new * p.num_kv_heads * p.head_dim * 2;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  return up(q) + up(k);
}

// The call as the attention path sees it once q and k_new have been rotated into (q_out, k_out):
// contiguous [B, S, H, D] tensors, no rotary.  Also used (with the original pointers) to size the
// workspace, so that the path chosen there is the path taken in launch_rope's caller.
vattn_fwd_params_t rope_rotated_view(const vattn_fwd_params_t& p, const void* q_out, const void* k_out) {
  vattn_fwd_params_t r = p;
  r.q = q_out;
  r.q_head_stride = p.head_dim;
  r.q_row_stride = (int64_t)p.num_heads * p.head_dim;
  r.q_batch_stride = (int64_t)p.seqlen_q * r.q_row_stride;
  r.k_new = k_out;
  r.knew_head_stride = p.head_dim;
  r.knew_row_stride = (int64_t)p.num_kv_heads * p.head_dim;
  r.knew_batch_stride = (int64_t)p.seqlen_new * r.knew_row_stride;
  r.rotary_cos = r.rotary_sin = nullptr;
  r.rotary_dim = 0;
  return r;
}

// Rotates q and k_new into `ws` and returns the parameters the attention path should run with.
vattn_fwd_params_t launch_rope(const vattn_fwd_params_t& p, cudaStream_t stream) {
  const size_t need = rope_workspace_bytes(p);
  if (!p.workspace || p.workspace_bytes < need)
    throw ArgError("[vattn] workspace too small: need " + std::to_string(need) + " bytes");
  const size_t q_bytes = ((size_t)p.batch * p.seqlen_q * p.num_heads * p.head_dim * 2 + 255) / 256 * 256;
  char* q_out = static_cast<char*>(p.workspace);
  char* k_out = q_out + q_bytes;
  const int64_t vecs =
      (int64_t)p.batch * (p.seqlen_q * p.num_heads + p.seqlen_new * p.num_kv_heads) * (p.head_dim / 8);
  int64_t blocks = (vecs + kThreads - 1) / kThreads;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  auto go = [&](auto tag) {
    using T = decltype(tag);
    rope_qk_kernel<T><<<(int)blocks, kThreads, 0, stream>>>(
        (const T*)p.q, p.q_batch_stride, p.q_row_stride, p.q_head_stride, (T*)q_out, (const T*)p.k_new,
        p.knew_batch_stride, p.knew_row_stride, p.knew_head_stride, (T*)k_out, (const T*)p.rotary_cos,
        (const T*)p.rotary_sin, p.cache_seqlens, p.batch, p.seqlen_q, p.seqlen_new, p.num_heads,
        p.num_kv_heads, p.head_dim, p.seqlen_k, p.rotary_dim, p.rotary_interleaved, p.causal);
  };
  if (p.dtype == VATTN_DTYPE_F16) go(__half{}); else go(__nv_bfloat16{});
  count_launch();
  VATTN_CUDA(cudaGetLastError());

  vattn_fwd_params_t r = rope_rotated_view(p, q_out, k_out);
  r.workspace = static_cast<char*>(p.workspace) + need;
  r.workspace_bytes = p.workspace_bytes - need;
  return r;
}

}  // namespace vattn
