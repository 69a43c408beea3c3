// C ABI, part B (attention operators): argument checks, kernel selection,
// workspace sizing.  Declarations: include/vattn_b200.h.
//
// Why decode QK^T / PV go to tensor cores when they can (VATTN_IMPL_TC): with a GQA
// group of g query heads every K/V byte needs 2*g FMAs (g=4: 2 FMA/B).  At the
// measured 6.57 TB/s that is 13 TFMA/s = 46-63 FMA/clk/SM, against a CUDA-core
// issue rate of 64 3-register FFMA/clk/SM (B300_MICROARCH.md: rt_SMSP = 2) before
// the bf16->fp32 conversions and the shuffle reductions.  The SIMT sweep is
// therefore instruction-bound, not HBM-bound, for g >= 2; the tcgen05 path puts
// keys on the MMA M axis so a 32 KB K tile costs 64 tensor cycles.
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "attn_common.cuh"
#include "capi_common.h"

namespace vattn {

std::atomic<uint64_t> g_launch_count{0};

void launch_cache_flat(const void*, const void*, void*, void*, int64_t, int64_t, int64_t, int64_t,
                       int64_t, int64_t, int, cudaStream_t);
void launch_append_kv(const vattn_fwd_params_t&, cudaStream_t);
void launch_simt(const vattn_fwd_params_t&, int, const SplitWorkspace&, cudaStream_t);
int simt_num_splits(const vattn_fwd_params_t&);
// tensor-core paths (attn_decode_tc.cu / attn_prefill_tc.cu)
bool decode_tc_supported(const vattn_fwd_params_t&, std::string* why);
bool decode_tc_fuses_append(const vattn_fwd_params_t&);
size_t decode_tc_workspace(const vattn_fwd_params_t&);
void launch_decode_tc(const vattn_fwd_params_t&, void* ws, size_t ws_bytes, cudaStream_t);
bool prefill_tc_supported(const vattn_fwd_params_t&, std::string* why);
size_t prefill_tc_workspace(const vattn_fwd_params_t&);
void launch_prefill_tc(const vattn_fwd_params_t&, void* ws, size_t ws_bytes, cudaStream_t);
int run_umma_selftest(char* buf, size_t len, cudaStream_t);
// fused prefill + decode (attn_pod_tc.cu)
bool pod_tc_supported(const vattn_fwd_params_t&, const vattn_fwd_params_t&, std::string* why);
size_t pod_tc_workspace(const vattn_fwd_params_t&, const vattn_fwd_params_t&);
// rope.cu
size_t rope_workspace_bytes(const vattn_fwd_params_t&);
vattn_fwd_params_t rope_rotated_view(const vattn_fwd_params_t&, const void*, const void*);
vattn_fwd_params_t launch_rope(const vattn_fwd_params_t&, cudaStream_t);
void launch_pod_tc(const vattn_fwd_params_t&, const vattn_fwd_params_t&, void* ws, size_t ws_bytes, cudaStream_t);
void launch_pod_dual(const vattn_fwd_params_t&, const vattn_fwd_params_t&, void* ws, size_t ws_bytes, cudaStream_t);

// ---- per-launch kernel timing -------------------------------------------------------
namespace {
struct TimingState {
  std::mutex mu;
  bool enabled = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;  // recycled event pairs
  size_t used = 0;
} g_timing;
}  // namespace


// ---- arrival counters (attn_common.cuh) ---------------------------------------------------------
namespace {
constexpr size_t kCounterSliceInts = 1 << 16;
constexpr int kCounterSlices = 32;
struct CounterPool {
  int* base = nullptr;
  int next = 0;
  std::unordered_map<uint64_t, int*> slices;
};
std::mutex g_counter_mu;
std::unordered_map<int, CounterPool> g_counter_pools;  // per device

int* zeroed_ints(size_t n) {
  // may run while the calling thread's stream is being captured: allocate and clear off to the side
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  VATTN_CUDA(cudaThreadExchangeStreamCaptureMode(&mode));
  int* ptr = nullptr;
  cudaError_t e = cudaMalloc(&ptr, n * sizeof(int));
  if (e == cudaSuccess) {
    cudaStream_t side;
    e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
      e = cudaMemsetAsync(ptr, 0, n * sizeof(int), side);
      if (e == cudaSuccess) e = cudaStreamSynchronize(side);
      cudaStreamDestroy(side);
    }
  }
  cudaThreadExchangeStreamCaptureMode(&mode);
  cuda_check(e, "arrival counter allocation");
  return ptr;
}
}  // namespace

int* arrival_counters(cudaStream_t stream, int user, size_t need) {
  int dev = 0;
  VATTN_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(g_counter_mu);
  CounterPool& pool = g_counter_pools[dev];
  const uint64_t key = (reinterpret_cast<uint64_t>(stream) << 2) ^ (uint64_t)user;
  auto it = pool.slices.find(key);
  if (it != pool.slices.end() && need <= kCounterSliceInts) return it->second;
  if (need > kCounterSliceInts || pool.next >= kCounterSlices) {
    // oversized request or more (stream, user) pairs than slices: a dedicated, never freed block
    int* big = zeroed_ints(need > kCounterSliceInts ? need : kCounterSliceInts);
    pool.slices[key] = big;
    return big;
  }
  if (!pool.base) pool.base = zeroed_ints(kCounterSliceInts * kCounterSlices);
  int* slice = pool.base + (size_t)pool.next++ * kCounterSliceInts;
  pool.slices[key] = slice;
  return slice;
}

int timing_begin(cudaStream_t stream) {
  std::lock_guard<std::mutex> g(g_timing.mu);
  if (!g_timing.enabled) return -1;
  if (g_timing.used == g_timing.pool.size()) {
    cudaEvent_t a, b;
    VATTN_CUDA(cudaEventCreate(&a));
    VATTN_CUDA(cudaEventCreate(&b));
    g_timing.pool.emplace_back(a, b);
  }
  const int slot = (int)g_timing.used++;
  VATTN_CUDA(cudaEventRecord(g_timing.pool[slot].first, stream));
  return slot;
}

void timing_end(int slot, cudaStream_t stream) {
  if (slot < 0) return;
  std::lock_guard<std::mutex> g(g_timing.mu);
  VATTN_CUDA(cudaEventRecord(g_timing.pool[slot].second, stream));
}

namespace {

int translate_attn_exception() {
  try {
    throw;
  } catch (const ArgError& e) {
    g_last_error = e.what();
    return VATTN_ERR_INVALID;
  } catch (const UnsupportedError& e) {
    g_last_error = e.what();
    return VATTN_ERR_UNSUPPORTED;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return VATTN_ERR_DRIVER;
  } catch (...) {
    g_last_error = "unknown C++ exception";
    return VATTN_ERR_DRIVER;
  }
}

void validate(const vattn_fwd_params_t& p) {
  if (!p.q || !p.k_cache || !p.v_cache || !p.out) throw ArgError("[vattn] null tensor pointer");
  if (p.dtype != VATTN_DTYPE_F16 && p.dtype != VATTN_DTYPE_BF16)
    throw ArgError("FlashAttention only support fp16 and bf16 data type");  // flash_api.cpp wording
  if (p.batch < 0 || p.seqlen_q <= 0 || p.seqlen_k < 0) throw ArgError("[vattn] bad sizes");
  if (p.num_heads <= 0 || p.num_kv_heads <= 0 || p.num_heads % p.num_kv_heads != 0)
    throw ArgError("Number of heads in key/value must divide number of heads in query");
  if (p.head_dim % 8 != 0) throw ArgError("head_size must be a multiple of 8");
  if ((p.k_new == nullptr) != (p.v_new == nullptr))
    throw ArgError("[vattn] k and v must be supplied together");
  if (p.k_new) {
    if (p.seqlen_new <= 0) throw ArgError("[vattn] seqlen_new must be positive when k is given");
    // flash_api.cpp:1454 -- the sarathi wrapper pattern-matches this text
    if (p.seqlen_new > p.seqlen_k)
      throw ArgError("If key is supplied, it must have seqlen <= the seqlen of the KV cache");
    // without cache_seqlens the append row would be seqlen_k itself: one row past the cache view
    // (possibly unmapped virtual memory)
    if (!p.cache_seqlens)
      throw ArgError("[vattn] cache_seqlens is required when k / v are appended");
    // the append / rotary kernels read k and v with 16-byte vector loads
    const int64_t ns[] = {p.knew_batch_stride, p.knew_row_stride, p.knew_head_stride,
                          p.vnew_batch_stride, p.vnew_row_stride, p.vnew_head_stride};
    for (int64_t s : ns)
      if (s % 8 != 0) throw ArgError("[vattn] strides must be multiples of 8 elements");
  }
  if ((p.rotary_cos == nullptr) != (p.rotary_sin == nullptr))
    throw ArgError("If rotary cos is provided, rotary sin must also be provided");  // flash_api.cpp:1516
  if (p.rotary_cos) {
    if (!p.k_new)  // flash_api.cpp:1504
      throw ArgError("If rotary cos/sin are provided, new key / value to be appended to KV cache must also be provided");
    if (p.rotary_dim <= 0 || p.rotary_dim > p.head_dim) throw ArgError("rotary_dim must be <= headdim");
    if (p.rotary_dim % 16 != 0)
      throw ArgError("Only rotary dimensions divisible by 16 are currently supported");
    if (p.seqlen_ro < p.seqlen_k) throw ArgError("cos/sin seqlen must be at least the seqlen of KV cache");
    if ((reinterpret_cast<uintptr_t>(p.rotary_cos) | reinterpret_cast<uintptr_t>(p.rotary_sin)) & 15)
      throw ArgError("[vattn] tensors must be 16-byte aligned");
  }
  auto al = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
  if (!al(p.q) || !al(p.k_cache) || !al(p.v_cache) || !al(p.out) || (p.k_new && !al(p.k_new)) ||
      (p.v_new && !al(p.v_new)))
    throw ArgError("[vattn] tensors must be 16-byte aligned");
  const int64_t st[] = {p.q_batch_stride, p.q_row_stride,  p.q_head_stride, p.k_batch_stride,
                        p.k_row_stride,   p.k_head_stride, p.v_batch_stride, p.v_row_stride,
                        p.v_head_stride,  p.o_batch_stride, p.o_row_stride,  p.o_head_stride};
  for (int64_t s : st)
    if (s % 8 != 0) throw ArgError("[vattn] strides must be multiples of 8 elements");
}

enum class Path { Simt, DecodeTc, PrefillTc };

Path choose(const vattn_fwd_params_t& p) {
  std::string why;
  if (p.impl == VATTN_IMPL_SIMT) return Path::Simt;
  const bool decode_like = p.seqlen_q == 1;
  if (decode_like) {
    if (decode_tc_supported(p, &why)) return Path::DecodeTc;
  } else {
    if (prefill_tc_supported(p, &why)) return Path::PrefillTc;
  }
  if (p.impl == VATTN_IMPL_TC)
    throw UnsupportedError("[vattn] tensor-core path not available for this call: " + why);
  return Path::Simt;
}

size_t workspace_for(const vattn_fwd_params_t& p, Path path) {
  switch (path) {
    case Path::DecodeTc: return decode_tc_workspace(p);
    case Path::PrefillTc: return prefill_tc_workspace(p);
    default: {
      const int s = simt_num_splits(p);
      return s > 1 ? split_workspace_bytes((int64_t)p.batch * p.seqlen_q, p.num_heads, s, p.head_dim)
                   : 0;
    }
  }
}

void run_fwd(const vattn_fwd_params_t& p_in, cudaStream_t stream) {
  validate(p_in);
  if (p_in.batch == 0) return;
  // rotary: q and k_new are rotated into the head of the workspace, everything below runs on those
  const vattn_fwd_params_t p = p_in.rotary_cos ? launch_rope(p_in, stream) : p_in;
  const Path path = choose(p);
  // the tensor-core decode kernel appends a single new token itself; every other case uses the
  // separate append kernel
  if (!(path == Path::DecodeTc && decode_tc_fuses_append(p))) launch_append_kv(p, stream);
  const size_t need = workspace_for(p, path);
  if (need > 0 && (!p.workspace || p.workspace_bytes < need))
    throw ArgError("[vattn] workspace too small: need " + std::to_string(need) + " bytes");
  switch (path) {
    case Path::DecodeTc: launch_decode_tc(p, p.workspace, p.workspace_bytes, stream); break;
    case Path::PrefillTc: launch_prefill_tc(p, p.workspace, p.workspace_bytes, stream); break;
    default: {
      const int s = simt_num_splits(p);
      SplitWorkspace ws{nullptr, nullptr};
      if (s > 1) ws = carve_workspace(p.workspace, (int64_t)p.batch * p.seqlen_q, p.num_heads, s, p.head_dim);
      launch_simt(p, s, ws, stream);
    }
  }
}

// device staging arena for the host-buffer entry point; one per caller stream: the lock only
// serialises the enqueue, so calls on two streams must not share staged q / k / out
struct HostStage {
  void* dev = nullptr;
  size_t cap = 0;
  void* ws = nullptr;
  size_t ws_cap = 0;
  void ensure(void** ptr, size_t* cap_, size_t bytes) {
    if (*cap_ >= bytes) return;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap_ = 0;
    VATTN_CUDA(cudaMalloc(ptr, bytes));
    *cap_ = bytes;
  }
};
std::mutex g_stage_mu;
std::unordered_map<cudaStream_t, HostStage> g_stages;

}  // namespace
}  // namespace vattn

using namespace vattn;

extern "C" {

size_t vattn_fwd_kvcache_workspace(const vattn_fwd_params_t* p) {
  if (!p) return 0;
  try {
    validate(*p);
    if (p->batch == 0) return 0;
    if (p->rotary_cos) {
      const vattn_fwd_params_t r = rope_rotated_view(*p, p->q, p->k_new);
      return rope_workspace_bytes(*p) + workspace_for(r, choose(r));
    }
    return workspace_for(*p, choose(*p));
  } catch (...) {
    translate_attn_exception();
    return 0;
  }
}

int vattn_fwd_kvcache(const vattn_fwd_params_t* p, void* stream) {
  if (!p) return VATTN_ERR_INVALID;
  try {
    run_fwd(*p, static_cast<cudaStream_t>(stream));
    return VATTN_OK;
  } catch (...) {
    return translate_attn_exception();
  }
}

int vattn_single_prefill(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                         int64_t k_row_stride, int64_t k_head_stride, const void* v,
                         int64_t v_row_stride, int64_t v_head_stride, void* out,
                         int64_t o_row_stride, int64_t o_head_stride, int32_t qo_len,
                         int32_t kv_len, int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                         int32_t dtype, int32_t causal, float sm_scale, void* workspace,
                         size_t workspace_bytes, void* stream) {
  // flashinfer.single_prefill_with_kv_cache == one request, no append, the whole k/v
  // visible, bottom-right aligned causal mask, scale 1/sqrt(D) by default
  // (vattention_flashinfer_wrapper.py:151-158 passes no scale).
  vattn_fwd_params_t p;
  std::memset(&p, 0, sizeof(p));
  p.q = q;
  p.q_batch_stride = 0, p.q_row_stride = q_row_stride, p.q_head_stride = q_head_stride;
  p.k_cache = const_cast<void*>(k);
  p.v_cache = const_cast<void*>(v);
  p.k_batch_stride = 0, p.k_row_stride = k_row_stride, p.k_head_stride = k_head_stride;
  p.v_batch_stride = 0, p.v_row_stride = v_row_stride, p.v_head_stride = v_head_stride;
  p.out = out;
  p.o_batch_stride = 0, p.o_row_stride = o_row_stride, p.o_head_stride = o_head_stride;
  p.batch = 1, p.cache_batch = 1, p.seqlen_q = qo_len, p.seqlen_k = kv_len;
  p.num_heads = num_heads, p.num_kv_heads = num_kv_heads, p.head_dim = head_dim;
  p.dtype = dtype, p.causal = causal;
  p.softmax_scale = sm_scale > 0.f ? sm_scale : 1.0f / sqrtf((float)head_dim);
  p.workspace = workspace, p.workspace_bytes = workspace_bytes;
  return vattn_fwd_kvcache(&p, stream);
}

static bool pod_fused_path(const vattn_fwd_params_t* prefill, const vattn_fwd_params_t* decode) {
  if (!prefill || !decode) return false;
  if (prefill->impl == VATTN_IMPL_SIMT || decode->impl == VATTN_IMPL_SIMT) return false;
  if (prefill->batch == 0 || decode->batch == 0) return false;
  std::string why;
  return pod_tc_supported(*prefill, *decode, &why);
}

size_t vattn_pod_workspace(const vattn_fwd_params_t* prefill, const vattn_fwd_params_t* decode) {
  try {
    if (prefill) validate(*prefill);
    if (decode) validate(*decode);
  } catch (...) {
    translate_attn_exception();
    return 0;
  }
  // enough for every strategy of vattn_pod_fwd: the persistent kernel or the two separate calls
  size_t a = prefill ? vattn_fwd_kvcache_workspace(prefill) : 0;
  size_t b = decode ? vattn_fwd_kvcache_workspace(decode) : 0;
  a = (a + 255) / 256 * 256;
  size_t fused = 0;
  try {
    if (pod_fused_path(prefill, decode)) fused = pod_tc_workspace(*prefill, *decode);
  } catch (...) {
    translate_attn_exception();
  }
  return a + b > fused ? a + b : fused;
}

namespace {
// side stream + fork/join events per caller stream for the co-scheduled POD strategy
struct PodSide {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
PodSide& pod_side_for(cudaStream_t main) {
  static std::mutex mu;
  static std::unordered_map<cudaStream_t, PodSide> sides;
  std::lock_guard<std::mutex> g(mu);
  PodSide& s = sides[main];
  if (!s.side) {
    VATTN_CUDA(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
    VATTN_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    VATTN_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  return s;
}
enum class PodStrategy { Streams, Kernel, Serial, Dual };
PodStrategy pod_strategy(int32_t fused_params) {
  // the reference: 15 = "pick the most suitable" (fused_api.cpp:24-53), anything else names one
  // tile configuration of its fused kernel.  Here:
  //   Dual    ONE launch, every CTA carries a prefill pipeline and a decode pipeline side by side
  //           (attn_pod_tc.cu: pod_dual_kernel) -- true sharing of an SM's tensor pipe and memory queue;
  //   Streams the two specialised kernels co-scheduled on two streams (fork / join inside the call);
  //   Kernel  the persistent one-item-at-a-time kernel (pod_tc_kernel);
  //   Serial  the two kernels back to back.
  // VATTN_POD_STRATEGY=dual|streams|kernel|serial overrides; otherwise 15 -> the default below, an
  // explicit configuration -> Kernel.
  if (const char* e = std::getenv("VATTN_POD_STRATEGY")) {
    const std::string v(e);
    if (v == "kernel") return PodStrategy::Kernel;
    if (v == "serial") return PodStrategy::Serial;
    if (v == "streams") return PodStrategy::Streams;
    if (v == "dual") return PodStrategy::Dual;
  }
  if (fused_params == 64) return PodStrategy::Dual;  // our own value: none of the reference's configurations
  return fused_params == 15 ? PodStrategy::Streams : PodStrategy::Kernel;
}

// fused_params == 15 ("pick the most suitable", fused_api.cpp:24-53): the dual-role kernel when the two
// sides are of comparable length, otherwise the two specialised kernels.  Time estimates from the
// shapes: prefill 4 Hq D Bp Sq Sk FLOP at ~0.9 PFLOP/s, decode 4 Hkv D Bd Sk bytes at ~6.9 TB/s.
// Measured (Llama-3-8B fp16, profiles/r2_pod_arms.jsonl): 1 x 2048 @ 16K + 64 x 16K decodes, ratio ~1:
// dual 1.081 ms vs 1.264 serial / 1.263 two streams / 1.359 persistent; 8 x 16K + 56 x 4K, ratio ~70:
// dual 20.1 ms vs 17.0 serial -- the decode side has nothing to hide behind there.
bool pod_auto_prefers_dual(const vattn_fwd_params_t& pre, const vattn_fwd_params_t& dec) {
  const double tp = 4.0 * pre.num_heads * pre.head_dim * (double)pre.batch * pre.seqlen_q * pre.seqlen_k / 0.9e15;
  const double td = 4.0 * dec.num_kv_heads * dec.head_dim * (double)dec.batch * dec.seqlen_k / 6.9e12;
  const double lo = tp < td ? tp : td, hi = tp < td ? td : tp;
  return hi > 0 && lo / hi >= 0.3;
}
}  // namespace

int vattn_pod_fwd(const vattn_fwd_params_t* prefill, const vattn_fwd_params_t* decode,
                  int32_t fused_params, void* workspace, size_t workspace_bytes, void* stream) {
  try {
    if (prefill) validate(*prefill);
    if (decode) validate(*decode);
    if ((prefill && prefill->rotary_cos) || (decode && decode->rotary_cos))
      throw UnsupportedError("[vattn] the fused POD call takes no rotary arguments (fused_attn_interface.py:12-40)");
    cudaStream_t main = static_cast<cudaStream_t>(stream);
    PodStrategy strat = pod_strategy(fused_params);
    if (fused_params == 15 && !std::getenv("VATTN_POD_STRATEGY") && pod_fused_path(prefill, decode) &&
        pod_auto_prefers_dual(*prefill, *decode))
      strat = PodStrategy::Dual;
    if (strat == PodStrategy::Kernel && pod_fused_path(prefill, decode)) {
      launch_pod_tc(*prefill, *decode, workspace, workspace_bytes, main);
      return VATTN_OK;
    }
    if (strat == PodStrategy::Dual && pod_fused_path(prefill, decode)) {
      launch_pod_dual(*prefill, *decode, workspace, workspace_bytes, main);
      return VATTN_OK;
    }
    // the two specialised kernels: co-scheduled on two streams, or back to back (one side missing,
    // serial requested)
    size_t a = prefill ? vattn_fwd_kvcache_workspace(prefill) : 0;
    a = (a + 255) / 256 * 256;
    const bool both = prefill && decode && prefill->batch > 0 && decode->batch > 0;
    const bool fork = both && strat != PodStrategy::Serial;
    PodSide* side = fork ? &pod_side_for(main) : nullptr;
    if (decode) {
      vattn_fwd_params_t d = *decode;
      d.workspace = workspace ? static_cast<char*>(workspace) + a : nullptr;
      d.workspace_bytes = workspace_bytes > a ? workspace_bytes - a : 0;
      if (fork) {
        VATTN_CUDA(cudaEventRecord(side->fork, main));
        VATTN_CUDA(cudaStreamWaitEvent(side->side, side->fork, 0));
        run_fwd(d, side->side);
        VATTN_CUDA(cudaEventRecord(side->join, side->side));
      } else if (!prefill) {
        run_fwd(d, main);
      }
    }
    if (prefill) {
      vattn_fwd_params_t p = *prefill;
      p.workspace = workspace;
      p.workspace_bytes = a;
      run_fwd(p, main);
    }
    if (fork) {
      VATTN_CUDA(cudaStreamWaitEvent(main, side->join, 0));
    } else if (decode && prefill) {
      vattn_fwd_params_t d = *decode;
      d.workspace = workspace ? static_cast<char*>(workspace) + a : nullptr;
      d.workspace_bytes = workspace_bytes > a ? workspace_bytes - a : 0;
      run_fwd(d, main);
    }
    return VATTN_OK;
  } catch (...) {
    return translate_attn_exception();
  }
}

int vattn_cache_flat(const void* key, const void* value, void* k_cache, void* v_cache,
                     int64_t num_tokens, int64_t row_elems, int64_t key_stride,
                     int64_t value_stride, int64_t k_cache_stride, int64_t v_cache_stride,
                     int32_t elem_bytes, void* stream) {
  try {
    if (num_tokens > 0 && (!key || !value || !k_cache || !v_cache)) throw ArgError("[vattn] null tensor pointer");
    launch_cache_flat(key, value, k_cache, v_cache, num_tokens, row_elems, key_stride, value_stride,
                      k_cache_stride, v_cache_stride, elem_bytes, static_cast<cudaStream_t>(stream));
    return VATTN_OK;
  } catch (...) {
    return translate_attn_exception();
  }
}

static int fwd_kvcache_host_impl(const vattn_fwd_params_t* hp, void* stream_, bool drain) {
  if (!hp) return VATTN_ERR_INVALID;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  try {
    if (hp->rotary_cos || hp->rotary_sin)
      throw UnsupportedError("[vattn] the host-buffer entry point takes no rotary tables (device-resident "
                             "cos/sin go through vattn_fwd_kvcache)");
    std::lock_guard<std::mutex> g(g_stage_mu);
    HostStage& g_stage = g_stages[stream];
    const size_t eb = 2;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t q_bytes = (size_t)hp->batch * hp->seqlen_q * hp->num_heads * hp->head_dim * eb;
    const size_t kn_bytes =
        hp->k_new ? (size_t)hp->batch * hp->seqlen_new * hp->num_kv_heads * hp->head_dim * eb : 0;
    const size_t idx_bytes = (size_t)hp->batch * sizeof(int32_t);
    const size_t total = up(q_bytes) * 2 + up(kn_bytes) * 2 + up(idx_bytes) * 2;
    g_stage.ensure(&g_stage.dev, &g_stage.cap, total);
    char* base = static_cast<char*>(g_stage.dev);
    char* d_q = base;
    char* d_o = d_q + up(q_bytes);
    char* d_kn = d_o + up(q_bytes);
    char* d_vn = d_kn + up(kn_bytes);
    char* d_sl = d_vn + up(kn_bytes);
    char* d_bi = d_sl + up(idx_bytes);

    vattn_fwd_params_t p = *hp;
    VATTN_CUDA(cudaMemcpyAsync(d_q, hp->q, q_bytes, cudaMemcpyHostToDevice, stream));
    p.q = d_q;
    p.q_head_stride = hp->head_dim;
    p.q_row_stride = (int64_t)hp->num_heads * hp->head_dim;
    p.q_batch_stride = p.q_row_stride * hp->seqlen_q;
    p.out = d_o;
    p.o_head_stride = p.q_head_stride, p.o_row_stride = p.q_row_stride, p.o_batch_stride = p.q_batch_stride;
    if (hp->k_new) {
      VATTN_CUDA(cudaMemcpyAsync(d_kn, hp->k_new, kn_bytes, cudaMemcpyHostToDevice, stream));
      VATTN_CUDA(cudaMemcpyAsync(d_vn, hp->v_new, kn_bytes, cudaMemcpyHostToDevice, stream));
      p.k_new = d_kn, p.v_new = d_vn;
      p.knew_head_stride = p.vnew_head_stride = hp->head_dim;
      p.knew_row_stride = p.vnew_row_stride = (int64_t)hp->num_kv_heads * hp->head_dim;
      p.knew_batch_stride = p.vnew_batch_stride = p.knew_row_stride * hp->seqlen_new;
    }
    if (hp->cache_seqlens) {
      VATTN_CUDA(cudaMemcpyAsync(d_sl, hp->cache_seqlens, idx_bytes, cudaMemcpyHostToDevice, stream));
      p.cache_seqlens = reinterpret_cast<const int32_t*>(d_sl);
    }
    if (hp->cache_batch_idx) {
      VATTN_CUDA(cudaMemcpyAsync(d_bi, hp->cache_batch_idx, idx_bytes, cudaMemcpyHostToDevice, stream));
      p.cache_batch_idx = reinterpret_cast<const int32_t*>(d_bi);
    }
    p.softmax_lse = nullptr;
    validate(p);
    const size_t need = p.batch ? workspace_for(p, choose(p)) : 0;
    if (need) g_stage.ensure(&g_stage.ws, &g_stage.ws_cap, need);
    p.workspace = g_stage.ws;
    p.workspace_bytes = g_stage.ws_cap;
    run_fwd(p, stream);
    VATTN_CUDA(cudaMemcpyAsync(hp->out, d_o, q_bytes, cudaMemcpyDeviceToHost, stream));
    if (drain) VATTN_CUDA(cudaStreamSynchronize(stream));
    return VATTN_OK;
  } catch (...) {
    return translate_attn_exception();
  }
}

int vattn_fwd_kvcache_host(const vattn_fwd_params_t* hp, void* stream) {
  return fwd_kvcache_host_impl(hp, stream, true);
}

int vattn_fwd_kvcache_host_async(const vattn_fwd_params_t* hp, void* stream) {
  return fwd_kvcache_host_impl(hp, stream, false);
}

// ---- pipelined host-buffer entry point ---------------------------------------------------------
// Same contract as vattn_fwd_kvcache_host_async, but the copies leave the compute stream: inputs of
// call i+1 travel host->device on a copy-in stream while the kernels of call i run, outputs travel
// device->host on a copy-out stream while the kernels of call i+1 run (two staging slots, events
// between the three streams).  The caller's stream therefore does NOT cover the last copies:
// vattn_host_pipeline_join(stream) makes it wait for every output still in flight -- call it once
// before synchronising (e.g. once per decode iteration).
namespace {
struct HostPipe {
  cudaStream_t in = nullptr, out = nullptr;
  cudaEvent_t h2d_done[2] = {nullptr, nullptr}, kernel_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  void* dev[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  void* ws = nullptr;
  size_t ws_cap = 0;
  uint64_t calls = 0;
};
std::mutex g_pipe_mu;
std::unordered_map<cudaStream_t, HostPipe> g_pipes;

HostPipe& pipe_for(cudaStream_t main) {
  HostPipe& hp = g_pipes[main];
  if (!hp.in) {
    VATTN_CUDA(cudaStreamCreateWithFlags(&hp.in, cudaStreamNonBlocking));
    VATTN_CUDA(cudaStreamCreateWithFlags(&hp.out, cudaStreamNonBlocking));
    for (int s = 0; s < 2; s++) {
      VATTN_CUDA(cudaEventCreateWithFlags(&hp.h2d_done[s], cudaEventDisableTiming));
      VATTN_CUDA(cudaEventCreateWithFlags(&hp.kernel_done[s], cudaEventDisableTiming));
      VATTN_CUDA(cudaEventCreateWithFlags(&hp.d2h_done[s], cudaEventDisableTiming));
    }
  }
  return hp;
}
void grow(void** ptr, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return;
  if (*ptr) VATTN_CUDA(cudaFree(*ptr));  // synchronises the device: nothing is using the old block after this
  *ptr = nullptr, *cap = 0;
  VATTN_CUDA(cudaMalloc(ptr, bytes));
  *cap = bytes;
}
}  // namespace

int vattn_fwd_kvcache_host_pipelined(const vattn_fwd_params_t* hp, void* stream_) {
  if (!hp) return VATTN_ERR_INVALID;
  cudaStream_t main = static_cast<cudaStream_t>(stream_);
  try {
    if (hp->rotary_cos || hp->rotary_sin)
      throw UnsupportedError("[vattn] the host-buffer entry points take no rotary tables");
    std::lock_guard<std::mutex> g(g_pipe_mu);
    HostPipe& pp = pipe_for(main);
    const int s = (int)(pp.calls & 1);
    const size_t eb = 2;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t q_bytes = (size_t)hp->batch * hp->seqlen_q * hp->num_heads * hp->head_dim * eb;
    const size_t kn_bytes =
        hp->k_new ? (size_t)hp->batch * hp->seqlen_new * hp->num_kv_heads * hp->head_dim * eb : 0;
    const size_t idx_bytes = (size_t)hp->batch * sizeof(int32_t);
    grow(&pp.dev[s], &pp.cap[s], up(q_bytes) * 2 + up(kn_bytes) * 2 + up(idx_bytes) * 2);
    char* d_q = static_cast<char*>(pp.dev[s]);
    char* d_o = d_q + up(q_bytes);
    char* d_kn = d_o + up(q_bytes);
    char* d_vn = d_kn + up(kn_bytes);
    char* d_sl = d_vn + up(kn_bytes);
    char* d_bi = d_sl + up(idx_bytes);

    // inputs: slot s was last read by the kernels of call i-2
    if (pp.used[s]) VATTN_CUDA(cudaStreamWaitEvent(pp.in, pp.kernel_done[s], 0));
    vattn_fwd_params_t p = *hp;
    VATTN_CUDA(cudaMemcpyAsync(d_q, hp->q, q_bytes, cudaMemcpyHostToDevice, pp.in));
    p.q = d_q;
    p.q_head_stride = hp->head_dim;
    p.q_row_stride = (int64_t)hp->num_heads * hp->head_dim;
    p.q_batch_stride = p.q_row_stride * hp->seqlen_q;
    p.out = d_o;
    p.o_head_stride = p.q_head_stride, p.o_row_stride = p.q_row_stride, p.o_batch_stride = p.q_batch_stride;
    if (hp->k_new) {
      VATTN_CUDA(cudaMemcpyAsync(d_kn, hp->k_new, kn_bytes, cudaMemcpyHostToDevice, pp.in));
      VATTN_CUDA(cudaMemcpyAsync(d_vn, hp->v_new, kn_bytes, cudaMemcpyHostToDevice, pp.in));
      p.k_new = d_kn, p.v_new = d_vn;
      p.knew_head_stride = p.vnew_head_stride = hp->head_dim;
      p.knew_row_stride = p.vnew_row_stride = (int64_t)hp->num_kv_heads * hp->head_dim;
      p.knew_batch_stride = p.vnew_batch_stride = p.knew_row_stride * hp->seqlen_new;
    }
    if (hp->cache_seqlens) {
      VATTN_CUDA(cudaMemcpyAsync(d_sl, hp->cache_seqlens, idx_bytes, cudaMemcpyHostToDevice, pp.in));
      p.cache_seqlens = reinterpret_cast<const int32_t*>(d_sl);
    }
    if (hp->cache_batch_idx) {
      VATTN_CUDA(cudaMemcpyAsync(d_bi, hp->cache_batch_idx, idx_bytes, cudaMemcpyHostToDevice, pp.in));
      p.cache_batch_idx = reinterpret_cast<const int32_t*>(d_bi);
    }
    VATTN_CUDA(cudaEventRecord(pp.h2d_done[s], pp.in));
    p.softmax_lse = nullptr;
    validate(p);
    const size_t need = p.batch ? workspace_for(p, choose(p)) : 0;
    if (need) grow(&pp.ws, &pp.ws_cap, need);  // one workspace: the kernels of successive calls are stream ordered
    p.workspace = pp.ws;
    p.workspace_bytes = pp.ws_cap;

    // kernels: after this call's inputs landed and after the output staging of call i-2 was read out
    VATTN_CUDA(cudaStreamWaitEvent(main, pp.h2d_done[s], 0));
    if (pp.used[s]) VATTN_CUDA(cudaStreamWaitEvent(main, pp.d2h_done[s], 0));
    run_fwd(p, main);
    VATTN_CUDA(cudaEventRecord(pp.kernel_done[s], main));

    // output: device -> host on the copy-out stream
    VATTN_CUDA(cudaStreamWaitEvent(pp.out, pp.kernel_done[s], 0));
    VATTN_CUDA(cudaMemcpyAsync(hp->out, d_o, q_bytes, cudaMemcpyDeviceToHost, pp.out));
    VATTN_CUDA(cudaEventRecord(pp.d2h_done[s], pp.out));
    pp.used[s] = true;
    pp.calls++;
    return VATTN_OK;
  } catch (...) {
    return translate_attn_exception();
  }
}

int vattn_host_pipeline_join(void* stream_) {
  cudaStream_t main = static_cast<cudaStream_t>(stream_);
  try {
    std::lock_guard<std::mutex> g(g_pipe_mu);
    auto it = g_pipes.find(main);
    if (it == g_pipes.end()) return VATTN_OK;
    for (int s = 0; s < 2; s++)
      if (it->second.used[s]) VATTN_CUDA(cudaStreamWaitEvent(main, it->second.d2h_done[s], 0));
    return VATTN_OK;
  } catch (...) {
    return translate_attn_exception();
  }
}

uint64_t vattn_launch_count(void) { return g_launch_count.load(); }

int vattn_kernel_timing(int op, double* total_ms, uint64_t* launches) {
  try {
    std::lock_guard<std::mutex> g(g_timing.mu);
    if (op == 1) {
      g_timing.enabled = true;
      g_timing.used = 0;
    } else if (op == 0) {
      g_timing.enabled = false;
      g_timing.used = 0;
    } else if (op == 2) {
      double sum = 0;
      for (size_t i = 0; i < g_timing.used; i++) {
        VATTN_CUDA(cudaEventSynchronize(g_timing.pool[i].second));
        float ms = 0;
        VATTN_CUDA(cudaEventElapsedTime(&ms, g_timing.pool[i].first, g_timing.pool[i].second));
        sum += ms;
      }
      if (total_ms) *total_ms = sum;
      if (launches) *launches = g_timing.used;
      g_timing.used = 0;
    } else {
      throw ArgError("[vattn] kernel_timing: op must be 0, 1 or 2");
    }
    return VATTN_OK;
  } catch (...) {
    return translate_attn_exception();
  }
}

int vattn_selftest_umma(char* buf, size_t buf_len, void* stream) {
  try {
    return run_umma_selftest(buf, buf_len, static_cast<cudaStream_t>(stream));
  } catch (...) {
    translate_attn_exception();
    if (buf && buf_len) {
      std::strncpy(buf, g_last_error.c_str(), buf_len - 1);
      buf[buf_len - 1] = 0;
    }
    return -1;
  }
}

}  // extern "C"
