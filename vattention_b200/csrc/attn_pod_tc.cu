// POD: prefill and decode attention in ONE persistent launch on sm_100a.
//
// Replaces pod_attn's fused kernel (true_fused_tb_fwd_kernel -> compute_fused_tb_attn,
// pod_attn/pod_attn/fused_fwd_kernel.h:1408-1590; launch fused_fwd_launch_template.h:350-458).
// The reference co-schedules small SM80 CTAs of both kinds: thread 0 of every CTA reads %smid,
// bumps tbAssign[sm] to pick prefill-vs-decode in a fixed ratio and a second counter to take the
// next tile of that kind (:1455-1491); the counters are cudaMalloc'ed and memset on every call
// (fused_fwd_launch_template.h:407-410).
//
// Here a tcgen05 prefill tile wants a whole SM (512 TMEM columns, 192 KB of staged tiles), so the
// fusion is expressed as one persistent CTA per SM that keeps taking work items from a single
// ticket counter; ticket t is a prefill item iff floor((t+1)P/T) > floor(tP/T) (P prefill items of
// T total), i.e. the two kinds are interleaved in proportion to their counts -- the reference's
// "proportional" policy (fused_params bit 0) -- so while some SMs run tensor-bound prefill tiles
// others stream K/V for decode chunks and both the tensor pipes and HBM stay busy for the whole
// launch.  The counter lives in the caller's workspace (one cudaMemsetAsync, no allocation).
// Work items are the same device functions the stand-alone kernels run (attn_tc_work.cuh), so
// each output equals the separate call's bit for bit (the reference asserts allclose(1e-3),
// pod_attn/tests/attn_sweep.py:82-97).
#include <climits>

#include "attn_common.cuh"
#include "attn_tc_host.h"
#include "attn_tc_work.cuh"
#include "sm100_ptx.cuh"

namespace vattn {

void launch_append_kv(const vattn_fwd_params_t&, cudaStream_t);

namespace {

using namespace ptx;
using namespace tcwork;

constexpr int kPodDecodeStages = 6;

struct __align__(1024) PodSmem {
  union {
    PrefillSmem prefill;     // chunks of <= 128 rows: one row block per item
    Prefill2Smem prefill2;   // otherwise two row blocks per item (ping-pong softmax warpgroups)
    DecodeSmemT<kPodDecodeStages> decode;
  } u;
  TcBarriers bar;
  uint32_t tmem_base;
  int ticket;
};

struct PodSched {
  int* counter;       // zeroed before launch
  long long n_prefill, n_decode;
  int prefill_blocks;  // row blocks per prefill item (1 or 2)
  int prefill_items_per_head;
};

template <typename T, int GP>
__global__ void __launch_bounds__(kPrefill2Threads, 1)
pod_tc_kernel(const __grid_constant__ CUtensorMap qmap_p, const __grid_constant__ CUtensorMap kmap_p,
              const __grid_constant__ CUtensorMap vmap_p, const __grid_constant__ CUtensorMap kmap_d,
              const __grid_constant__ CUtensorMap vmap_d, const __grid_constant__ CUtensorMap kt_p,
              const __grid_constant__ CUtensorMap vt_p, const __grid_constant__ CUtensorMap kt_d,
              const __grid_constant__ CUtensorMap vt_d, const PrefillParams pp, const DecodeTcParams dp,
              const PodSched sch) {
  extern __shared__ uint8_t smem_raw[];
  PodSmem& sm = *reinterpret_cast<PodSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    prefetch_tensormap(&qmap_p);
    prefetch_tensormap(&kmap_p);
    prefetch_tensormap(&vmap_p);
    prefetch_tensormap(&kmap_d);
    prefetch_tensormap(&vmap_d);
    // every barrier word holds a live mbarrier from here on, so work items can inval + re-init
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
  const uint32_t tmem = sm.tmem_base;
  const long long total = sch.n_prefill + sch.n_decode;

  for (;;) {
    if (threadIdx.x == 0) sm.ticket = atomicAdd(sch.counter, 1);
    __syncthreads();
    const long long t = sm.ticket;
    __syncthreads();  // everyone has read the ticket before thread 0 may overwrite it
    if (t >= total) break;
    const long long np0 = t * sch.n_prefill / total, np1 = (t + 1) * sch.n_prefill / total;
    if (np1 > np0) {
      // prefill item np0, in the stand-alone kernel's launch order: row blocks of one (batch, head)
      // are consecutive (heavy first), so the CTAs running at the same time share that head's
      // K/V in L2 instead of streaming 148 different heads through it
      const int mi = sch.prefill_items_per_head - 1 - (int)(np0 % sch.prefill_items_per_head);
      const long long rem = np0 / sch.prefill_items_per_head;
      const int h = (int)(rem % pp.num_heads), b = (int)(rem / pp.num_heads);
      if (sch.prefill_blocks == 2) {
        PrefillSegment seg{};
        seg.mt2 = mi, seg.h = h, seg.b = b, seg.j0 = 0, seg.j1 = INT_MAX, seg.parts = 1;
        prefill2_work<T>(&qmap_p, &kmap_p, &vmap_p, &kt_p, &vt_p, pp, sm.u.prefill2, sm.bar, tmem, seg, true);
      }
      else
        prefill_work<T>(&qmap_p, &kmap_p, &vmap_p, &kt_p, &vt_p, pp, sm.u.prefill, sm.bar, tmem, mi, h, b, true);
    } else {
      const long long d = t - np0;  // decode item: chunk fastest, then kv head, then batch
      const int chunk = (int)(d % dp.num_chunks);
      const long long r = d / dp.num_chunks;
      decode_work<T, GP, kPodDecodeStages>(&kmap_d, &vmap_d, &kt_d, &vt_d, dp, sm.u.decode, sm.bar, tmem, chunk,
                                           (int)(r % dp.num_kv_heads), (int)(r / dp.num_kv_heads), true);
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---- dual-role CTA: a prefill pipeline and a decode pipeline resident on EVERY SM at the same time ----
// The reference's point is SM-level co-residency: a compute-bound prefill CTA and a memory-bound
// decode CTA share one SM's tensor pipe and its memory queue (fused_fwd_kernel.h:1455-1491).  A
// tcgen05 kernel cannot get there with two kernels (a full-size prefill CTA owns the SM's shared
// memory and TMEM), so ONE persistent CTA per SM carries both: warps {0, 1, 4-7} are a prefill
// pipeline (TMA producer, MMA issuer, softmax warpgroup; one 128-row block, Q 32 KB + 3 x 32 KB ring,
// TMEM S x2 with P in place + O = 384 columns) and warps {2, 3, 8-11} a decode pipeline (2 x 32 KB
// ring, 64 TMEM columns).  Each side synchronises on its own named barrier and takes its own items
// from its own ticket counter, so the decode side streams K/V continuously while the prefill side
// keeps the tensor pipe busy -- neither waits for the other until both queues are empty.
constexpr int kDualPrefillStages = 3, kDualDecodeStages = 2;

struct __align__(1024) PodDualSmem {
  PrefillSmemT<kDualPrefillStages> prefill;
  DecodeSmemT<kDualDecodeStages> decode;
  TcBarriers bar_p, bar_d;
  uint32_t tmem_base;
  int ticket_p, ticket_d;
};

template <typename T, int GP>
__global__ void __launch_bounds__(kPrefill2Threads, 1)
pod_dual_kernel(const __grid_constant__ CUtensorMap qmap_p, const __grid_constant__ CUtensorMap kmap_p,
                const __grid_constant__ CUtensorMap vmap_p, const __grid_constant__ CUtensorMap kmap_d,
                const __grid_constant__ CUtensorMap vmap_d, const __grid_constant__ CUtensorMap kt_p,
                const __grid_constant__ CUtensorMap vt_p, const __grid_constant__ CUtensorMap kt_d,
                const __grid_constant__ CUtensorMap vt_d, const PrefillParams pp, const DecodeTcParams dp,
                const PodSched sch) {
  extern __shared__ uint8_t smem_raw[];
  PodDualSmem& sm = *reinterpret_cast<PodDualSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    prefetch_tensormap(&qmap_p);
    prefetch_tensormap(&kmap_p);
    prefetch_tensormap(&vmap_p);
    prefetch_tensormap(&kmap_d);
    prefetch_tensormap(&vmap_d);
    TcBarriers* both[2] = {&sm.bar_p, &sm.bar_d};
    for (TcBarriers* bar : both) {  // live barriers from here on: work items inval + re-init
      for (int s = 0; s < kMaxStages; s++) {
        mbar_init(&bar->full[s], 1);
        mbar_init(&bar->empty[s], 1);
      }
      mbar_init(&bar->q_full, 1);
      for (int i = 0; i < 2; i++) {
        mbar_init(&bar->s_full[i], 1);
        mbar_init(&bar->p_ready[i], 1);
        mbar_init(&bar->o_full[i], 1);
      }
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
  const uint32_t tmem = sm.tmem_base;
  const bool prefill_side = warp < 2 || (warp >= 4 && warp < 8);
  if (prefill_side) {
    Side sd;
    sd.tma_warp = 0, sd.mma_warp = 1, sd.sm_warp0 = 4, sd.bar_id = 4, sd.nthreads = 192;
    for (;;) {
      if (sd.tid() == 0) sm.ticket_p = atomicAdd(sch.counter, 1);
      sd.sync();
      const long long t = sm.ticket_p;
      sd.sync();  // everyone has read the ticket before it is overwritten
      if (t >= sch.n_prefill) break;
      // the stand-alone kernel's order: row blocks of one (batch, head) consecutive, heavy first
      const int mi = sch.prefill_items_per_head - 1 - (int)(t % sch.prefill_items_per_head);
      const long long rem = t / sch.prefill_items_per_head;
      const int h = (int)(rem % pp.num_heads), b = (int)(rem / pp.num_heads);
      // S0 | S1 at columns [0, 256) (P in place), O at [256, 384)
      prefill_work<T, kDualPrefillStages, true>(&qmap_p, &kmap_p, &vmap_p, &kt_p, &vt_p, pp, sm.prefill, sm.bar_p,
                                                tmem, mi, h, b, true, tmem + 256, sd);
    }
  } else {
    Side sd;
    sd.tma_warp = 2, sd.mma_warp = 3, sd.sm_warp0 = 8, sd.bar_id = 5, sd.nthreads = 192;
    for (;;) {
      if (sd.tid() == 0) sm.ticket_d = atomicAdd(sch.counter + 1, 1);
      sd.sync();
      const long long d = sm.ticket_d;
      sd.sync();
      if (d >= sch.n_decode) break;
      const int chunk = (int)(d % dp.num_chunks);  // chunk fastest, then kv head, then batch
      const long long r = d / dp.num_chunks;
      decode_work<T, GP, kDualDecodeStages>(&kmap_d, &vmap_d, &kt_d, &vt_d, dp, sm.decode, sm.bar_d, tmem + 384,
                                            chunk, (int)(r % dp.num_kv_heads), (int)(r / dp.num_kv_heads), true, sd);
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <typename T>
void launch_pod_dual_t(const vattn_fwd_params_t& pre, const vattn_fwd_params_t& dec, void* ws, cudaStream_t stream) {
  // workspace: [2 int counters | pad to 256][decode split partials]
  int* counter = static_cast<int*>(ws);
  void* dec_ws = static_cast<char*>(ws) + 256;
  PrefillTcLaunch P;
  DecodeTcLaunch Dl;
  build_prefill_tc(pre, &P);
  build_decode_tc(dec, dec_ws, stream, &Dl, false);
  PodSched sch;
  sch.counter = counter;
  sch.prefill_blocks = 1;
  sch.prefill_items_per_head = P.pp.num_m_tiles;
  sch.n_prefill = (long long)P.pp.num_m_tiles * pre.num_heads * pre.batch;
  sch.n_decode = (long long)Dl.dp.num_chunks * dec.num_kv_heads * dec.batch;
  VATTN_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int), stream));
  if (!decode_tc_fuses_append(dec)) launch_append_kv(dec, stream);
  const int sms = num_sms();
  const long long most = sch.n_prefill > sch.n_decode ? sch.n_prefill : sch.n_decode;
  const int grid = (int)(most < sms ? most : sms);
  const size_t smem = sizeof(PodDualSmem) + 1024;
  const int group = dec.num_heads / dec.num_kv_heads;
  auto launch = [&](auto kernel) {
    VATTN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tslot = timing_begin(stream);
    kernel<<<grid, kPrefill2Threads, smem, stream>>>(P.qmap, P.kmap, P.vmap, Dl.kmap, Dl.vmap, P.kmap_tail,
                                                     P.vmap_tail, Dl.kmap_tail, Dl.vmap_tail, P.pp, Dl.dp, sch);
    timing_end(tslot, stream);
  };
  if (group <= 4) launch(pod_dual_kernel<T, 4>);
  else if (group <= 8) launch(pod_dual_kernel<T, 8>);
  else launch(pod_dual_kernel<T, 16>);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
  if (!Dl.dp.arrive && Dl.dp.num_chunks > 1) launch_combine(dec, Dl.dp.num_chunks, Dl.ws, stream);
}

template <typename T>
void launch_pod_t(const vattn_fwd_params_t& pre, const vattn_fwd_params_t& dec, void* ws, cudaStream_t stream) {
  // workspace: [int counter | pad to 256][decode split partials]
  int* counter = static_cast<int*>(ws);
  void* dec_ws = static_cast<char*>(ws) + 256;
  PrefillTcLaunch P;
  DecodeTcLaunch Dl;
  build_prefill_tc(pre, &P);
  build_decode_tc(dec, dec_ws, stream, &Dl, false);
  PodSched sch;
  sch.counter = counter;
  sch.prefill_blocks = pre.seqlen_q > kTile ? 2 : 1;
  sch.prefill_items_per_head = (P.pp.num_m_tiles + sch.prefill_blocks - 1) / sch.prefill_blocks;
  sch.n_prefill = (long long)sch.prefill_items_per_head * pre.num_heads * pre.batch;
  sch.n_decode = (long long)Dl.dp.num_chunks * dec.num_kv_heads * dec.batch;
  VATTN_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
  if (!decode_tc_fuses_append(dec)) launch_append_kv(dec, stream);
  const int sms = num_sms();
  const long long total = sch.n_prefill + sch.n_decode;
  const int grid = (int)(total < sms ? total : sms);
  const size_t smem = sizeof(PodSmem) + 1024;
  const int group = dec.num_heads / dec.num_kv_heads;
  auto launch = [&](auto kernel) {
    VATTN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tslot = timing_begin(stream);
    kernel<<<grid, kPrefill2Threads, smem, stream>>>(P.qmap, P.kmap, P.vmap, Dl.kmap, Dl.vmap, P.kmap_tail,
                                                     P.vmap_tail, Dl.kmap_tail, Dl.vmap_tail, P.pp, Dl.dp, sch);
    timing_end(tslot, stream);
  };
  if (group <= 4) launch(pod_tc_kernel<T, 4>);
  else if (group <= 8) launch(pod_tc_kernel<T, 8>);
  else launch(pod_tc_kernel<T, 16>);
  count_launch();
  VATTN_CUDA(cudaGetLastError());
  if (!Dl.dp.arrive && Dl.dp.num_chunks > 1) launch_combine(dec, Dl.dp.num_chunks, Dl.ws, stream);
}

}  // namespace

bool pod_tc_supported(const vattn_fwd_params_t& pre, const vattn_fwd_params_t& dec, std::string* why) {
  if (pre.dtype != dec.dtype) {
    if (why) *why = "prefill and decode dtypes differ";
    return false;
  }
  if (pre.seqlen_q < 2 || dec.seqlen_q != 1) {
    if (why) *why = "expects seqlen_q > 1 on the prefill side and == 1 on the decode side";
    return false;
  }
  if (pre.k_new) {
    if (why) *why = "append on the prefill side is not part of the fused call";
    return false;
  }
  return prefill_tc_supported(pre, why) && decode_tc_supported(dec, why);
}

size_t pod_tc_workspace(const vattn_fwd_params_t&, const vattn_fwd_params_t& dec) {
  return 256 + decode_tc_workspace_grid(dec);
}

void launch_pod_dual(const vattn_fwd_params_t& pre, const vattn_fwd_params_t& dec, void* ws, size_t ws_bytes,
                     cudaStream_t stream) {
  if (!ws || ws_bytes < pod_tc_workspace(pre, dec)) throw ArgError("[vattn] POD workspace too small");
  if (pre.dtype == VATTN_DTYPE_BF16) launch_pod_dual_t<__nv_bfloat16>(pre, dec, ws, stream);
  else launch_pod_dual_t<__half>(pre, dec, ws, stream);
}

void launch_pod_tc(const vattn_fwd_params_t& pre, const vattn_fwd_params_t& dec, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
  if (!ws || ws_bytes < pod_tc_workspace(pre, dec)) throw ArgError("[vattn] POD workspace too small");
  if (pre.dtype == VATTN_DTYPE_BF16) launch_pod_t<__nv_bfloat16>(pre, dec, ws, stream);
  else launch_pod_t<__half>(pre, dec, ws, stream);
}

}  // namespace vattn
