/*
 * vattn_b200.h -- C ABI of the B200-native vAttention hot path.
 *
 * One shared library (libvattn_b200.so) exports everything below with C linkage,
 * plain pointers and sizes only (no torch / ATen / pybind types).  It is what a
 * maintainer of the reference would bind from its own extension layer:
 *
 *   part A  KV-cache allocator  -> replaces vattention/apis.h:1-63 (the 13 pybind
 *           names registered at vattention/vattention.cu:614-637) and the VMM
 *           backend in vattention/cudaInternal.h:15-94, vtensor.h:21-46.
 *   part B  attention operators -> replaces the third-party kernels the sarathi
 *           wrappers dispatch to:
 *             flash_attn_with_kvcache        vattention_flashattention_wrapper.py:159-166,194-205
 *             single_prefill_with_kv_cache   vattention_flashinfer_wrapper.py:151-158
 *             true_fused_attn_with_kvcache   vattention_flashattention_pod_wrapper.py:177-191
 *                                            (pod_attn/pod_attn/fused_attn_interface.py:12-137)
 *             cache_flat                     sarathi-lean/csrc/cache_kernels.cu:524-570
 *
 * Conventions
 *   - every entry point that can fail returns an int status: 0 = ok, < 0 = error;
 *     vattn_last_error() returns the message of the last failure on the calling
 *     thread (the text a Python binding should raise as RuntimeError).
 *   - device pointers are raw CUDA virtual addresses; strides are in ELEMENTS;
 *     the innermost (head_dim) dimension must be contiguous, exactly as the
 *     reference requires (fused_attn_interface.py:80-81).
 *   - `stream` is a CUstream / cudaStream_t passed as void*; NULL = legacy default.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef VATTN_B200_H_
#define VATTN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VATTN_OK 0
#define VATTN_ERR_INVALID (-1)  /* bad argument / bad configuration            */
#define VATTN_ERR_OOM (-2)      /* "OOM on demand" / "page pool is empty"       */
#define VATTN_ERR_DRIVER (-3)   /* a CUDA driver / runtime call failed          */
#define VATTN_ERR_STATE (-4)    /* call sequence error (e.g. not initialised)   */
#define VATTN_ERR_UNSUPPORTED (-5)

/* message of the last failure on this thread ("" if none) */
const char* vattn_last_error(void);
/* library version string, e.g. "vattn_b200 0.1 (sm_100a)" */
const char* vattn_version(void);

/* ------------------------------------------------------------------------ */
/* Part A: virtually-contiguous KV-cache allocator                           */
/* ------------------------------------------------------------------------ */

typedef struct vattn_allocator vattn_allocator_t; /* opaque */

/* VMM backends.  CUDA = cuMemAddressReserve/cuMemCreate/cuMemMap through
 * libcuda (resolved at run time, so the library loads on a box without a
 * driver).  HOST_MOCK = same bookkeeping against a recording fake driver (no
 * GPU); exists so the page-map arithmetic can be checked bit-exactly on CPU.
 * It never backs a tensor that a kernel touches. */
#define VATTN_BACKEND_CUDA 0
#define VATTN_BACKEND_HOST_MOCK 1

int vattn_create(vattn_allocator_t** out, int backend);
int vattn_destroy(vattn_allocator_t* a);

/* Derived configuration (vattention/vattention.cu:38-74 arithmetic). */
typedef struct {
  uint64_t num_layers, num_kv_heads, head_size, max_batch_size, max_context_length;
  uint64_t bytes_per_elem, page_size, megacache;
  uint64_t tokens_per_page;          /* vattention.cu:41-44   */
  uint64_t virt_buff_size_per_token; /* vattention.cu:53-56   */
  uint64_t virt_buff_size_per_req;   /* vattention.cu:57-67   */
  uint64_t virt_buff_size;           /* vattention.cu:69      */
  uint64_t max_pages_per_req;        /* vattention.cu:60      */
  uint64_t phys_granularity;         /* what the driver reported */
  uint64_t num_tensors;              /* 2*L, or 2 with megacache */
} vattn_config_t;

/* apis.h:3-13 init_kvcache.  Reserves VA only.  On return ptrs[0..n) holds the
 * base device address of [K_0..K_{L-1}, V_0..V_{L-1}] (or [K, V] with
 * megacache) -- the order of the tensor list the reference returns
 * (vattention.cu:163-186).  `ptrs` must have room for 2*num_layers entries.
 * shape/ndim describe one tensor: {B, maxlen, Hkv, D} or {B, maxlen, L, Hkv, D};
 * strides are contiguous (vtensor.h:101-102).  page_size must be a multiple of
 * the device's minimum VMM granularity (2 MiB on B200).                      */
int vattn_init_kvcache(vattn_allocator_t* a, uint64_t num_layers, uint64_t num_kv_heads,
                       uint64_t head_size, uint64_t max_batch_size,
                       uint64_t max_context_length, int device, uint64_t bytes_per_elem,
                       uint64_t page_size, int megacache, uint64_t* ptrs, int* n_ptrs,
                       int64_t shape[5], int* ndim);
int vattn_get_config(vattn_allocator_t* a, vattn_config_t* out);

/* apis.h:23-25 reserve_physical_pages: pre-creates free_memory/page_size handles
 * rounded down to a multiple of 2*L (utils.h:221-228); returns the pool size,
 * or < 0 on error.                                                            */
int64_t vattn_reserve_physical_pages(vattn_allocator_t* a, uint64_t free_memory);

/* apis.h:27-29 step (all mapping on the critical path). seq_lens has n ==
 * max_batch_size entries; 0 marks an inactive slot.                           */
int vattn_step(vattn_allocator_t* a, const uint64_t* seq_lens, size_t n, int eager_reclaim);
/* apis.h:31-35 step_async: maps what THIS step needs before returning, then
 * hands "what step+1 .. step+9 will need" to the persistent mapper thread.   */
int vattn_step_async(vattn_allocator_t* a, const uint64_t* seq_lens, size_t n);
/* apis.h:53-59 */
int vattn_alloc_new_batch_idx(vattn_allocator_t* a, uint64_t seqlen); /* reqId, or -1 */
int vattn_free_batch_idx(vattn_allocator_t* a, int req_id);
/* apis.h:61-63 */
uint64_t vattn_num_free_kvblocks(vattn_allocator_t* a);
/* apis.h:41-43: unmap everything, free VA, release handles */
int vattn_cleanup(vattn_allocator_t* a);
/* apis.h:37-39, 45-47, 15-21 */
void vattn_set_verbose(vattn_allocator_t* a, int on);
void vattn_set_deferred_reclamation(vattn_allocator_t* a, int on);
void vattn_show_kvcache_config(vattn_allocator_t* a);
void vattn_show_allocator_state(vattn_allocator_t* a);
/* apis.h:49-51 map_common_pages (prefix-sharing proof of concept) */
int vattn_map_common_pages(vattn_allocator_t* a, uint64_t num_tokens);

/* --- B200-native additions (no reference equivalent) --- */
/* Block until the mapper thread is idle (the reference spins on an atomic at
 * the top of the NEXT step_async, utils.h:160-164).                          */
int vattn_wait_background(vattn_allocator_t* a);
/* Register the stream attention kernels run on.  Before any cuMemUnmap the
 * allocator waits for an event recorded on this stream at the last step call,
 * so a page is never pulled from under an in-flight kernel.  enable = 0 turns
 * the fence off (stream NULL with enable = 1 means the legacy default stream). */
int vattn_set_compute_stream(vattn_allocator_t* a, void* stream, int enable);
/* step_async may be QUEUED behind the mapper pass still in flight instead of waiting for it, when
 * the new lengths need no page beyond those every request already held when that pass started and
 * the pass cannot take pages back: the mapper then runs exactly the operations the reference runs
 * (pass, prepare = no-op, pass) while the caller goes on launching kernels.  on = 0 restores the
 * reference's wait at the top of every step_async (utils.h:160-164).  Default: on.              */
int vattn_set_queueing(vattn_allocator_t* a, int on);
/* Timing of the last step / background pass, nanoseconds (host clock), and totals since
 * init_kvcache.  The call waits for the mapper to be idle (like every call but step_async): a loop
 * that must not stand still reads the totals once at its end.                                   */
typedef struct {
  uint64_t critical_path_ns;   /* time inside the last step / step_async call  */
  uint64_t background_ns;      /* duration of the last mapper-thread pass      */
  uint64_t sync_pages_mapped;  /* pages (not blocks) mapped on the critical path */
  uint64_t async_pages_mapped; /* pages mapped by the last background pass     */
  uint64_t driver_calls;       /* cumulative driver VMM calls                  */
  uint64_t total_critical_path_ns, total_background_ns, max_background_ns;
  uint64_t total_sync_pages, total_async_pages;
  uint64_t steps, passes;      /* step / step_async calls; mapper passes       */
  uint64_t queued_steps;       /* step_async calls that rode behind a pass     */
} vattn_step_stats_t;
int vattn_get_step_stats(vattn_allocator_t* a, vattn_step_stats_t* out);

/* --- introspection used by the parity tests --- */
/* copies mapped_pages[] and curr_seq_lengths[] (utils.h:68-69) */
int vattn_get_state(vattn_allocator_t* a, uint64_t* mapped_pages, uint64_t* seq_lens, size_t n);
/* free pool, bottom -> top (the reference pops from the back, mux.h:1-8). Each
 * id is the 0-based creation index of the handle.  Returns the pool size.    */
size_t vattn_get_free_pool(vattn_allocator_t* a, uint64_t* ids, size_t cap);
/* page map (utils.h:24-29), sorted by (reqId, offset, layer); 5 words per
 * entry: reqId, req_offset, layer, k_id, v_id.  Returns the entry count.    */
size_t vattn_get_pagemap(vattn_allocator_t* a, uint64_t* words, size_t cap_entries);
/* HOST_MOCK only: driver-call log, 4 words per record: op, va, size, handle id
 * (op: 1 reserve, 2 create, 3 map, 4 set_access, 5 unmap, 6 release, 7 addr_free) */
size_t vattn_get_driver_log(vattn_allocator_t* a, uint64_t* words, size_t cap_records);
void vattn_clear_driver_log(vattn_allocator_t* a);
/* HOST_MOCK only: the mock driver's "device" holds `bytes` of physical memory: a create that
 * would exceed it fails like cuMemCreate does when the device is full (0 = unlimited).       */
void vattn_mock_set_capacity(vattn_allocator_t* a, uint64_t bytes);
/* HOST_MOCK only: every map / set_access / unmap of the mock driver takes `us` microseconds (a slow
 * driver: several processes mapping at once), so host tests can show what the caller waits for.  */
void vattn_mock_set_call_delay_us(vattn_allocator_t* a, uint64_t us);
/* HOST_MOCK only: fence records / host waits seen by the mock driver, per slot {rec0, rec1, wait0, wait1} */
void vattn_mock_fence_counts(vattn_allocator_t* a, uint64_t out[4]);

/* ------------------------------------------------------------------------ */
/* Part B: attention over the contiguous K/V                                 */
/* ------------------------------------------------------------------------ */

#define VATTN_DTYPE_F16 0
#define VATTN_DTYPE_BF16 1

/* kernel selection for vattn_fwd_kvcache (0 = pick automatically) */
#define VATTN_IMPL_AUTO 0
#define VATTN_IMPL_SIMT 1 /* 128-bit vectorised sweep, warp-reduce softmax       */
#define VATTN_IMPL_TC 2   /* TMA + tcgen05 (sm_100a tensor cores, TMEM)          */

/*
 * flash_attn_with_kvcache semantics (arithmetic: pod_attn/pod_attn/flash_api.cpp:1291-1580,
 * block_info.h:11-44, mask.h:172, softmax.h:66-160, flash_fwd_kernel.h:685-790):
 *   for b in [0,batch):  slot = cache_batch_idx ? cache_batch_idx[b] : b
 *     L0 = cache_seqlens ? cache_seqlens[b] : seqlen_k
 *     if k_new: rows [L0, L0+seqlen_new) of slot <- k_new[b], v_new[b]
 *     Lk = L0 + seqlen_new
 *     out[b,i,h] = softmax_j(scale * q[b,i,h].k[slot,j,h/g]) . v[slot,j,h/g]
 *        over j < Lk, and when causal also j <= i + Lk - seqlen_q.
 *     a query row with no visible key produces zeros.
 */
typedef struct {
  /* q [batch, seqlen_q, num_heads, head_dim] */
  const void* q;
  int64_t q_batch_stride, q_row_stride, q_head_stride;
  /* caches [cache_batch, seqlen_k, num_kv_heads, head_dim]; rows >= the mapped
   * prefix of a slot may be unmapped VA and are never touched               */
  void* k_cache;
  void* v_cache;
  int64_t k_batch_stride, k_row_stride, k_head_stride;
  int64_t v_batch_stride, v_row_stride, v_head_stride;
  /* optional append [batch, seqlen_new, num_kv_heads, head_dim] (NULL = none) */
  const void* k_new;
  const void* v_new;
  int64_t knew_batch_stride, knew_row_stride, knew_head_stride;
  int64_t vnew_batch_stride, vnew_row_stride, vnew_head_stride;
  /* out [batch, seqlen_q, num_heads, head_dim] */
  void* out;
  int64_t o_batch_stride, o_row_stride, o_head_stride;
  /* optional fp32 log-sum-exp [batch, num_heads, seqlen_q] (NULL = none) */
  float* softmax_lse;
  const int32_t* cache_seqlens;   /* [batch] or NULL */
  const int32_t* cache_batch_idx; /* [batch] or NULL */
  int32_t batch, cache_batch, seqlen_q, seqlen_k, seqlen_new;
  int32_t num_heads, num_kv_heads, head_dim;
  int32_t dtype;  /* VATTN_DTYPE_* */
  int32_t causal; /* 0 / 1 */
  float softmax_scale;
  int32_t impl;       /* VATTN_IMPL_* */
  int32_t num_splits; /* 0 = heuristic */
  /* scratch for split-KV partials; size from vattn_fwd_kvcache_workspace().
   * May be NULL when that returns 0.                                        */
  void* workspace;
  size_t workspace_bytes;
  /* Rotary embedding applied to q and k_new before the append (flash_api.cpp:1298-1310,1503-1527;
   * flash_fwd_kernel.h:684-830).  cos/sin: [seqlen_ro, rotary_dim/2] contiguous, dtype of q;
   * NULL = none.  Requires k_new; rotary_dim % 16 == 0, <= head_dim; seqlen_ro >= seqlen_k.
   * rotary_interleaved: 1 pairs dims (2j, 2j+1), 0 pairs (j, j + rotary_dim/2).             */
  const void* rotary_cos;
  const void* rotary_sin;
  int32_t rotary_dim, rotary_interleaved, seqlen_ro;
} vattn_fwd_params_t;

size_t vattn_fwd_kvcache_workspace(const vattn_fwd_params_t* p);
int vattn_fwd_kvcache(const vattn_fwd_params_t* p, void* stream);

/* flashinfer.single_prefill_with_kv_cache(q[c,Hq,D], k[n,Hkv,D], v[n,Hkv,D],
 * causal) (vattention_flashinfer_wrapper.py:151-158): one request, NHD layout,
 * scale = 1/sqrt(D) unless sm_scale > 0.  Strides in elements.              */
int vattn_single_prefill(const void* q, int64_t q_row_stride, int64_t q_head_stride,
                         const void* k, int64_t k_row_stride, int64_t k_head_stride,
                         const void* v, int64_t v_row_stride, int64_t v_head_stride,
                         void* out, int64_t o_row_stride, int64_t o_head_stride,
                         int32_t qo_len, int32_t kv_len, int32_t num_heads,
                         int32_t num_kv_heads, int32_t head_dim, int32_t dtype,
                         int32_t causal, float sm_scale, void* workspace,
                         size_t workspace_bytes, void* stream);

/* POD: one launch computes a prefill problem and a decode problem
 * (true_fused_attn_with_kvcache, fused_attn_interface.py:12-137; host
 * fused_api.cpp:282-488).  Either side may be NULL (degenerates to the other,
 * fused_attn_interface.py:40-78).  fused_params is accepted for signature
 * compatibility (8/9/10/11/15/64, fused_api.cpp:24-53); scheduling here is a
 * persistent per-SM work queue, see DESIGN.md.                              */
size_t vattn_pod_workspace(const vattn_fwd_params_t* prefill, const vattn_fwd_params_t* decode);
int vattn_pod_fwd(const vattn_fwd_params_t* prefill, const vattn_fwd_params_t* decode,
                  int32_t fused_params, void* workspace, size_t workspace_bytes, void* stream);

/* cache_flat(key[c,Hkv,D], value, k_cache[>=c,Hkv,D], v_cache, "auto")
 * (cache_kernels.cu:524-570): k_cache[t,:,:] = key[t,:,:] for t < num_tokens.
 * Row strides in elements; each row is num_kv_heads*head_dim contiguous
 * elements (the reference indexes i < H*D off the row base, :505-518).      */
int vattn_cache_flat(const void* key, const void* value, void* k_cache, void* v_cache,
                     int64_t num_tokens, int64_t row_elems, int64_t key_stride,
                     int64_t value_stride, int64_t k_cache_stride, int64_t v_cache_stride,
                     int32_t elem_bytes, void* stream);

/* Reference-facing call with HOST buffers (bench.py's `e2e` leg): q / k_new /
 * v_new / cache_seqlens / cache_batch_idx / out in `p` are HOST pointers
 * (pinned for full speed); the caches stay device pointers (they are the
 * resident state the allocator owns).  Copies in, runs vattn_fwd_kvcache,
 * copies `out` back, all on `stream`; returns after the stream is drained.
 * Host tensors must be densely packed ([batch, seqlen, heads, dim]).        */
int vattn_fwd_kvcache_host(const vattn_fwd_params_t* p, void* stream);
/* Same, but returns as soon as the copies and kernels are enqueued: the result is in `out` once
 * `stream` has been synchronised.  Staging is stream-ordered, so consecutive calls on one stream
 * (the layers of a decode iteration) need no host synchronisation in between; the host buffers
 * must stay untouched until the stream is drained.                                            */
int vattn_fwd_kvcache_host_async(const vattn_fwd_params_t* p, void* stream);

/* Pipelined variant: the host->device copies of call i+1 and the device->host copy of call i-1
 * overlap the kernels of call i (own copy streams, two staging slots).  `stream` alone no longer
 * covers the output copies: call vattn_host_pipeline_join(stream) once before synchronising it
 * (e.g. once per decode iteration); host buffers must stay untouched until then.               */
int vattn_fwd_kvcache_host_pipelined(const vattn_fwd_params_t* p, void* stream);
int vattn_host_pipeline_join(void* stream);

/* One-shot all-reduce(sum) over NVLink peer memory for the head-sharded attention block: the
 * only collective on the path (the o_proj output all-reduce, tensor_parallel/layers.py:448-451 ->
 * mappings.py:16-26, NCCL in the reference).  peer_partial_ptrs[r] / peer_flag_ptrs[r] are the
 * addresses, in THIS process, of rank r's partial buffer [n_elems] and flag array [world] inside a
 * symmetric (peer-mapped) allocation; the caller alternates two such buffer/flag sets on
 * consecutive calls and passes a strictly increasing `epoch` (same on all ranks).  Writes the sum
 * to `out` (local).  See csrc/tp_allreduce.cu for the protocol. */
int vattn_allreduce_oneshot(const uint64_t* peer_partial_ptrs, const uint64_t* peer_flag_ptrs, void* out,
                            int64_t n_elems, int dtype, int rank, int world, uint32_t epoch, void* stream);

/* The same collective fused with the GEMM that feeds it (the row-parallel o_proj,
 * tensor_parallel/layers.py:432-461): out[tokens, hidden] = sum over ranks of
 * x_r[tokens, k_local] . w_r[hidden, k_local]^T, ONE kernel per rank (csrc/oproj_allreduce.cu).
 * x: this rank's attention output (row stride in elements), w: this rank's o_proj weight shard in
 * nn.Linear layout [hidden, k_local] row-major.  peer_recv_ptrs[r] / peer_flag_ptrs[r]: rank r's
 * receive area (vattn_oproj_allreduce_recv_bytes) and flag array (..._flag_bytes, zeroed once) in a
 * symmetric allocation, as mapped in THIS process.  epoch_state: 4 zero-initialised uint32 in local
 * device memory ([0] completed calls, [2] != 0 after a peer failed to arrive); the call sequence
 * must be the same on every rank.  1..128 tokens, hidden % 32 == 0, k_local % 64 == 0;
 * CUDA-graph capturable (no host-side epoch).                                                     */
size_t vattn_oproj_allreduce_recv_bytes(int32_t max_tokens, int32_t hidden, int32_t world);
size_t vattn_oproj_allreduce_flag_bytes(int32_t hidden);
int vattn_oproj_allreduce(const void* x, int64_t x_row_stride, const void* w, void* out, int32_t tokens,
                          int32_t hidden, int32_t k_local, int32_t dtype, int32_t max_tokens,
                          const uint64_t* peer_recv_ptrs, const uint64_t* peer_flag_ptrs,
                          uint32_t* epoch_state, int32_t rank, int32_t world, void* stream);

/* number of kernel launches issued by this library since load (bench.py's
 * `gpu_launches` counts from here) */
uint64_t vattn_launch_count(void);

/* Per-launch device timing of the DOMINANT kernel of each operator call (the attention
 * sweep itself, not the append / combine helpers), taken with CUDA events on the
 * launching stream.  op: 1 = enable, 0 = disable and clear, 2 = read: waits for the
 * recorded events, returns the sum of elapsed milliseconds and the number of launches
 * since the last read, then clears.  bench.py derives `roofline.achieved` from it. */
int vattn_kernel_timing(int op, double* total_ms, uint64_t* launches);

/* Device-side self tests of the tcgen05/TMA building blocks (used by
 * tests/ on the GPU box): returns 0 when every variant matches the host
 * reference, else the index (1-based) of the first failing variant; writes a
 * human-readable report into buf.                                           */
int vattn_selftest_umma(char* buf, size_t buf_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VATTN_B200_H_ */
