// KvAllocator -- virtually-contiguous KV-cache allocator (host C++ over the
// CUDA driver VMM API).  Capability parity with the reference's
// vAttentionCachingAllocator (vattention/vattention.cu:27-610); every policy
// function cites the reference lines whose observable bookkeeping it must
// reproduce bit-exactly (mapped_pages[], free-pool order, page map,
// num_free_kvblocks, reqId choice).
//
// What is deliberately different (B200-first, see DESIGN.md "Allocator"):
//   * one persistent mapper thread parked on a condition variable instead of a
//     detached std::thread per step (vattention.cu:538-546); the context is made
//     current on it once;
//   * state lives in an object, not file-scope globals (utils.h:12-81); while a pass is
//     in flight the mapper thread owns the bookkeeping and every API call first waits
//     for it to be idle, so none of the reference's racy reads (SURVEY 5 hazards a-c)
//     exist.  One exception, step_async itself: when the lengths of the next step need
//     no page the in-flight pass has not already been seen to hold (and that pass
//     cannot take pages back), the step is QUEUED behind the pass instead of waiting
//     for it -- the mapper runs the same operations in the same order the reference
//     would (pass, no-op prepare, pass), only the caller does not stand still
//     meanwhile.  A slow driver (eight processes mapping at once) then costs nothing
//     on the critical path as long as eager mapping stays one page ahead;
//   * cuMemSetAccess is issued once per contiguous range per tensor instead of
//     once per page (cudaInternal.h:77-80);
//   * cuMemUnmap is fenced behind an event recorded on the compute stream;
//   * driver errors raise instead of exit(1) (cudaInternal.h:1-13).
#pragma once
#include <condition_variable>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <mutex>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "vmm_driver.h"

namespace vattn {

struct OomError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct InvalidError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct StateError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct KvConfig {
  u64 num_layers = 0, num_kv_heads = 0, head_size = 0, max_batch_size = 0;
  u64 max_context_length = 0, bytes_per_elem = 0, page_size = 0;
  bool megacache = false;
  int device = 0;
  u64 tokens_per_page = 0, per_token = 0, per_req = 0, virt_size = 0;
  u64 max_pages_per_req = 0, granularity = 0;
  // page_size < granularity: pages are a LOGICAL bookkeeping unit (the reference's 64/128/256 KB UVM
  // modes); physical memory is mapped in granularity-sized chunks, each backing `phys_group`
  // consecutive logical pages of one request in one tensor
  u64 phys_group = 1;
};

struct StepStats {
  u64 critical_path_ns = 0, background_ns = 0;      // last step call / last background pass
  u64 sync_pages_mapped = 0, async_pages_mapped = 0;
  // since init_kvcache (a caller that does not want to wait for the mapper every step reads these once)
  u64 total_critical_path_ns = 0, total_background_ns = 0;
  u64 total_sync_pages = 0, total_async_pages = 0;
  u64 max_background_ns = 0, steps = 0, passes = 0, queued_steps = 0;
};

struct PhysPage {
  u64 handle = 0;  // driver handle
  u64 id = 0;      // 0-based creation index (what parity tests compare)
};

class KvAllocator {
 public:
  explicit KvAllocator(std::unique_ptr<VmmDriver> drv);
  ~KvAllocator();

  std::vector<u64> init_kvcache(u64 num_layers, u64 num_kv_heads, u64 head_size,
                                u64 max_batch_size, u64 max_context_length, int device,
                                u64 bytes_per_elem, u64 page_size, bool megacache);
  u64 reserve_physical_pages(u64 free_memory);
  void step_sync(const u64* seq_lens, size_t n, bool eager_reclaim);
  void step_async(const u64* seq_lens, size_t n);
  int alloc_new_batch_idx(u64 seqlen);
  void free_batch_idx(int req_id);
  u64 num_free_kvblocks();
  void cleanup();
  void set_verbose(bool v);
  void set_deferred_reclamation(bool v);
  void show_kvcache_config();
  void show_allocator_state();
  void map_common_pages(u64 num_tokens);

  void wait_background();
  void set_compute_stream(void* stream, bool enable);
  void set_queueing(bool on);  // step_async may ride behind an in-flight pass (default on)
  StepStats stats();
  KvConfig config();
  void get_state(u64* mapped, u64* lens, size_t n);
  std::vector<u64> free_pool_ids();
  std::vector<u64> pagemap_words();
  VmmDriver* driver() { return drv_.get(); }

 private:
  using Key = std::tuple<u64, u64, u64>;  // (reqId, req_offset, layer)  utils.h:24

  // ---- policy (called by an API thread with mu_ held and the mapper idle, or by the mapper thread
  // during its pass: busy_ hands it the bookkeeping, so it runs WITHOUT mu_) ----
  u64 tokens_to_pages(u64 t) const { return (t + cfg_.tokens_per_page - 1) / cfg_.tokens_per_page; }
  u64 blocks_in_pool() const;
  bool kvblocks_available(u64 n) const { return blocks_in_pool() >= n; }
  u64 overcommitted() const;
  PhysPage pop_page();
  void map_pair(u64 req, u64 layer, u64 off, PhysPage k, PhysPage v);
  bool logical() const { return cfg_.phys_group > 1; }
  void chunk_ref(u64 tensor_base, u64 off);    // logical mode: back the chunk holding `off`
  void chunk_unref(u64 tensor_base, u64 off);  // ... and drop it when its last page goes
  void grow(u64 req, u64 nblocks, bool sync, u64* pages_counter);
  void unmap_one(u64 req);
  void release_some(u64 req, u64 retain);
  void map_for_curr_step(u64 req, u64 seq_len, u64* pages_counter);
  void reclaim_on_demand(u64 nblocks);
  void do_reclaim_pages();
  u64 need_new_page_async(u64 req, u64 eager) const;
  bool pass_will_reclaim() const;
  void background_pass();
  void hand_off_locked();                                    // snapshot + wake the mapper
  bool can_ride_behind(const u64* seq_lens, size_t n) const;  // see step_async
  void dump_state_locked();
  void log(const std::string& s) const;
  void require_configured() const;

  void mapper_main();
  // waits (with lk held) until the mapper has finished the pass it was given
  void wait_idle(std::unique_lock<std::mutex>& lk);

  std::unique_ptr<VmmDriver> drv_;
  KvConfig cfg_;
  bool configured_ = false;
  bool verbose_ = false;
  bool deferred_reclaim_ = true;  // utils.h:78

  std::vector<u64> k_ptr_, v_ptr_;  // VA bases, one per layer (1 with megacache)
  std::vector<PhysPage> pool_;      // free pages; back() is popped first (mux.h:1-8)
  u64 created_ = 0;
  std::map<Key, std::pair<PhysPage, PhysPage>> pagemap_;
  std::unordered_map<u64, u64> shared_refs_;  // page id -> live mappings (map_common_pages)
  // logical mode: physical chunk pool and, per chunk VA, (handle, number of logical pages inside)
  std::vector<u64> chunk_pool_;
  u64 slack_chunks_ = 0;  // chunks created on demand for partly used last chunks (<= B x tensors)
  std::unordered_map<u64, std::pair<u64, u64>> chunks_;
  std::vector<u64> mapped_pages_, seq_lens_;

  void* compute_stream_ = nullptr;
  bool fence_enabled_ = false;
  int fence_slot_ = 0;  // the driver fence slot of the step the mapper is (or was last) working on
  StepStats stats_;

  // what the in-flight pass started from (written under mu_ at hand-off, read under mu_): lower
  // bounds an API thread may rely on while the mapper owns the live bookkeeping
  std::vector<u64> handoff_lens_, handoff_mapped_;
  bool handoff_reclaims_ = false;  // that pass takes pages back (reclaim_on_demand)
  bool queueing_ = true;
  bool queued_ = false;            // one step rides behind the in-flight pass
  std::vector<u64> queued_lens_;

  std::mutex mu_;
  std::condition_variable cv_;
  std::thread mapper_;
  bool job_pending_ = false, busy_ = false, stop_ = false;
  std::string bg_error_;
};

}  // namespace vattn
