// See kv_allocator.h.  Reference: vattention/vattention.cu, utils.h, mux.h.
#include "kv_allocator.h"

#include <chrono>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace vattn {

namespace {
inline u64 now_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}
constexpr u64 kEagerNumSteps = 10;    // vattention.cu:486
constexpr u64 kEagerNumKvBlocks = 2;  // vattention.cu:487
const char* kOomMsg = "***** OOM on demand: not enough free pages to continue *****";
}  // namespace

KvAllocator::KvAllocator(std::unique_ptr<VmmDriver> drv) : drv_(std::move(drv)) {}

KvAllocator::~KvAllocator() {
  {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !busy_; });
    stop_ = true;
  }
  cv_.notify_all();
  if (mapper_.joinable()) mapper_.join();
}

void KvAllocator::log(const std::string& s) const {
  if (verbose_) std::cout << s << std::endl;  // utils.h:230-238
}

void KvAllocator::require_configured() const {
  if (!configured_) throw StateError("[vattn] init_kvcache has not been called");
}

void KvAllocator::wait_idle(std::unique_lock<std::mutex>& lk) {
  cv_.wait(lk, [&] { return !busy_; });
  if (!bg_error_.empty()) {
    std::string e;
    e.swap(bg_error_);
    throw std::runtime_error("[vattn] mapper thread failed: " + e);
  }
}

// ------------------------------------------------------------------ init ---

std::vector<u64> KvAllocator::init_kvcache(u64 num_layers, u64 num_kv_heads, u64 head_size,
                                           u64 max_batch_size, u64 max_context_length, int device,
                                           u64 bytes_per_elem, u64 page_size, bool megacache) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  if (configured_) throw StateError("[vattn] init_kvcache called twice without cleanup()");
  // vattention.cu:107-110 (asserts there; errors here) and utils.h:99-103
  if (!(max_batch_size > 0 && max_batch_size < 1000))
    throw InvalidError("[vattn] max_batch_size must be in (0, 1000)");
  if (!(max_context_length > 0 && max_context_length < 1000000))
    throw InvalidError("[vattn] max_context_length must be in (0, 1000000)");
  if (!(num_layers > 0 && num_layers < 100)) throw InvalidError("[vattn] num_layers must be in (0, 100)");
  if (!(num_kv_heads > 0 && num_kv_heads < 256))
    throw InvalidError("[vattn] num_kv_heads must be in (0, 256)");
  if (head_size == 0 || bytes_per_elem == 0 || page_size == 0)
    throw InvalidError("[vattn] head_size, bytes_per_elem and page_size must be non-zero");

  KvConfig c;
  c.num_layers = num_layers;
  c.num_kv_heads = num_kv_heads;
  c.head_size = head_size;
  c.max_batch_size = max_batch_size;
  c.max_context_length = max_context_length;
  c.bytes_per_elem = bytes_per_elem;
  c.page_size = page_size;
  c.megacache = megacache;
  c.device = device;

  // vattention.cu:38-47 (do_cuda_init -> granularity).  The reference asserts
  // granularity == page_size (cudaInternal.h:33); any multiple is mappable.
  c.granularity = drv_->init(device);
  if (page_size < c.granularity) {
    // the reference's UVM modes (utils.h:83-86, uvmInternal.h:219-226): 64/128/256 KB pages need its
    // patched nvidia-uvm driver.  Here they are a logical unit: bookkeeping (tokens_per_page, pool,
    // num_free_kvblocks) follows the reference formulas exactly, physical memory is mapped in
    // granularity-sized chunks
    if (c.granularity % page_size != 0)
      throw InvalidError("[vattn] page_size " + std::to_string(page_size) +
                         " must divide the device VMM granularity " + std::to_string(c.granularity));
    c.phys_group = c.granularity / page_size;
  } else if (page_size % c.granularity != 0) {
    throw InvalidError("[vattn] page_size " + std::to_string(page_size) +
                       " is not a multiple of the device VMM granularity " +
                       std::to_string(c.granularity));
  }

  // vattention.cu:41-44, 53-56
  c.per_token = num_kv_heads * head_size * bytes_per_elem * (megacache ? num_layers : 1);
  c.tokens_per_page = page_size / c.per_token;
  if (c.tokens_per_page == 0)
    throw InvalidError("[vattn] page_size is smaller than one token's K (or V) row");
  // vattention.cu:57-67
  u64 raw = c.per_token * max_context_length;
  c.per_req = (raw + page_size - 1) / page_size * page_size;
  c.max_pages_per_req = c.per_req / page_size;
  c.virt_size = c.per_req * max_batch_size;
  // vtensor.h:73-88: the tensor the caller sees has stride(0) == raw bytes, the
  // mapper uses per_req; they must agree or request r's pages land at the wrong
  // rows.  The reference throws this message when the total is not a multiple of
  // page*B; we also reject the cases its rounding lets through silently.
  if (raw % page_size != 0)
    throw InvalidError("size_bytes is not a multiple of page_size * shape[0]");
  if (c.phys_group > 1 && c.per_req % c.granularity != 0)
    throw InvalidError("[vattn] logical pages: per-request bytes (" + std::to_string(c.per_req) +
                       ") must be a multiple of the device VMM granularity so that a physical chunk "
                       "never spans two requests");

  u64 nt = megacache ? 1 : num_layers;
  std::vector<u64> k(nt), v(nt);
  // vattention.cu:163-186: K tensors are reserved first, then V.
  const u64 align = page_size < c.granularity ? c.granularity : page_size;
  for (u64 i = 0; i < nt; i++) k[i] = drv_->reserve(c.virt_size, align);
  for (u64 i = 0; i < nt; i++) v[i] = drv_->reserve(c.virt_size, align);

  cfg_ = c;
  k_ptr_ = k;
  v_ptr_ = v;
  mapped_pages_.assign(max_batch_size, 0);  // utils.h:88-97
  seq_lens_.assign(max_batch_size, 0);
  pagemap_.clear();
  shared_refs_.clear();
  chunks_.clear();
  stats_ = StepStats{};
  queued_ = false;
  configured_ = true;
  log("Initialized CUDA context and memory config etc...");
  log("num_tokens_per_kvblock: " + std::to_string(c.tokens_per_page));
  if (!mapper_.joinable()) mapper_ = std::thread([this] { mapper_main(); });

  std::vector<u64> out;
  out.insert(out.end(), k.begin(), k.end());
  out.insert(out.end(), v.begin(), v.end());
  return out;
}

u64 KvAllocator::reserve_physical_pages(u64 free_memory) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  // utils.h:221-228 + cudaInternal.h:45-59: the loop bound is the FREE-pool size
  u64 n = free_memory / cfg_.page_size;
  n -= n % (2 * cfg_.num_layers);
  log("Reserving " + std::to_string(n) + " pages of size " + std::to_string(cfg_.page_size) + " ...");
  if (!logical()) {
    while (pool_.size() < n) {
      PhysPage p;
      p.handle = drv_->create(cfg_.page_size);
      p.id = created_++;
      pool_.push_back(p);
    }
    return pool_.size();
  }
  // Logical (sub-granularity) pages: physical memory moves in 2 MiB chunks.  The chunks that back
  // the new logical bytes are created FIRST -- if the driver runs out, nothing has been published
  // and the call can be retried -- and total exactly free_memory (rounded up to one chunk).  The
  // extra chunk a request needs when the last chunk of one of its tensors is only partly used is
  // created on demand (chunk_ref), at most max_batch_size x tensors of them over the allocator's
  // life: the budget is exceeded only by fragmentation that actually occurs, never up front.
  if (pool_.size() < n) {
    const u64 want = (n * cfg_.page_size + cfg_.granularity - 1) / cfg_.granularity;
    const size_t had = chunk_pool_.size();
    try {
      for (u64 have = chunk_pool_.size() + chunks_.size() - slack_chunks_; have < want; have++)
        chunk_pool_.push_back(drv_->create(cfg_.granularity));
    } catch (...) {
      while (chunk_pool_.size() > had) {
        drv_->release(chunk_pool_.back());
        chunk_pool_.pop_back();
      }
      throw;
    }
    while (pool_.size() < n) {
      PhysPage p;
      p.handle = 0;
      p.id = created_++;
      pool_.push_back(p);
    }
  }
  return pool_.size();
}

// ---------------------------------------------------------------- policy ---

u64 KvAllocator::blocks_in_pool() const {
  // utils.h:8-11
  return cfg_.megacache ? pool_.size() / 2 : pool_.size() / (2 * cfg_.num_layers);
}

u64 KvAllocator::overcommitted() const {
  // utils.h:177-183 -- u64 arithmetic, wraps exactly like the reference
  u64 acc = 0;
  for (u64 r = 0; r < cfg_.max_batch_size; r++) acc += mapped_pages_[r] - tokens_to_pages(seq_lens_[r]);
  return acc;
}

PhysPage KvAllocator::pop_page() {
  // mux.h:1-8
  if (pool_.empty()) throw OomError("***** page pool is empty *****");
  PhysPage p = pool_.back();
  pool_.pop_back();
  return p;
}

void KvAllocator::map_pair(u64 req, u64 layer, u64 off, PhysPage k, PhysPage v) {
  // cudaInternal.h:70-82 minus the per-page cuMemSetAccess (batched by callers)
  // either both halves are mapped and recorded, or neither
  if (logical()) {
    chunk_ref(k_ptr_[layer], off);
    try {
      chunk_ref(v_ptr_[layer], off);
    } catch (...) {
      chunk_unref(k_ptr_[layer], off);
      throw;
    }
  } else {
    drv_->map(k_ptr_[layer] + off, cfg_.page_size, k.handle);
    try {
      drv_->map(v_ptr_[layer] + off, cfg_.page_size, v.handle);
    } catch (...) {
      drv_->unmap(k_ptr_[layer] + off, cfg_.page_size);
      throw;
    }
  }
  pagemap_[Key(req, off, layer)] = std::make_pair(k, v);
}

void KvAllocator::chunk_ref(u64 base, u64 off) {
  const u64 va = base + off / cfg_.granularity * cfg_.granularity;
  auto it = chunks_.find(va);
  if (it != chunks_.end()) {
    it->second.second++;
    return;
  }
  if (chunk_pool_.empty()) {
    // every budgeted chunk is in use and this (request, tensor) starts a new, partly used one
    const u64 tensors = 2 * (cfg_.megacache ? 1 : cfg_.num_layers);
    if (slack_chunks_ >= cfg_.max_batch_size * tensors) throw OomError("***** page pool is empty *****");
    chunk_pool_.push_back(drv_->create(cfg_.granularity));  // may throw: nothing has been changed yet
    slack_chunks_++;
  }
  const u64 h = chunk_pool_.back();
  drv_->map(va, cfg_.granularity, h);
  try {
    drv_->set_access(va, cfg_.granularity);
  } catch (...) {
    drv_->unmap(va, cfg_.granularity);
    throw;
  }
  chunk_pool_.pop_back();
  chunks_[va] = std::make_pair(h, (u64)1);
}

void KvAllocator::chunk_unref(u64 base, u64 off) {
  const u64 va = base + off / cfg_.granularity * cfg_.granularity;
  auto it = chunks_.find(va);
  if (it == chunks_.end()) throw StateError("[vattn] physical chunk missing on unmap");
  if (--it->second.second > 0) return;
  drv_->unmap(va, cfg_.granularity);
  chunk_pool_.push_back(it->second.first);
  chunks_.erase(it);
}

void KvAllocator::grow(u64 req, u64 nblocks, bool sync, u64* pages_counter) {
  // vattention.cu:268-323
  if (nblocks == 0) return;
  if (!kvblocks_available(nblocks)) {
    if (!sync) return;  // background attempts are best effort
    verbose_ = true;
    log("free pages: " + std::to_string(blocks_in_pool()));
    log("required: " + std::to_string(nblocks));
    dump_state_locked();
    throw OomError(kOomMsg);
  }
  const u64 nl = cfg_.megacache ? 1 : cfg_.num_layers;
  const u64 base = req * cfg_.per_req;
  const u64 first_off = base + mapped_pages_[req] * cfg_.page_size;
  u64 done = 0;
  for (u64 count = 0; count < nblocks; count++) {
    u64 off = base + mapped_pages_[req] * cfg_.page_size;  // utils.h:185-191
    if (!(off < (req + 1) * cfg_.per_req)) break;          // vattention.cu:254-266
    u64 layer = 0;
    try {
      for (; layer < nl; layer++) {
        PhysPage k = pop_page();  // K first, then V (mux.h:40-42)
        PhysPage v;
        try {
          v = pop_page();
        } catch (...) {
          pool_.push_back(k);
          throw;
        }
        try {
          map_pair(req, layer, off, k, v);
        } catch (...) {
          pool_.push_back(v);  // restore the LIFO order: K was popped first
          pool_.push_back(k);
          throw;
        }
      }
    } catch (...) {
      // undo the layers of this block that were already mapped (reverse order), then let the
      // blocks completed before it stand: page map, pool and mapped_pages_ stay consistent
      while (layer-- > 0) {
        auto it = pagemap_.find(Key(req, off, layer));
        if (logical()) {
          chunk_unref(k_ptr_[layer], off);
          chunk_unref(v_ptr_[layer], off);
        } else {
          drv_->unmap(k_ptr_[layer] + off, cfg_.page_size);
          drv_->unmap(v_ptr_[layer] + off, cfg_.page_size);
        }
        pool_.push_back(it->second.second);
        pool_.push_back(it->second.first);
        pagemap_.erase(it);
      }
      if (done) {
        for (u64 l2 = 0; l2 < nl && !logical(); l2++) {
          drv_->set_access(k_ptr_[l2] + first_off, done * cfg_.page_size);
          drv_->set_access(v_ptr_[l2] + first_off, done * cfg_.page_size);
        }
        if (pages_counter) *pages_counter += done * 2 * nl;
      }
      throw;
    }
    mapped_pages_[req]++;
    done++;
  }
  if (done) {
    for (u64 layer = 0; layer < nl && !logical(); layer++) {  // (logical mode grants access per chunk)
      drv_->set_access(k_ptr_[layer] + first_off, done * cfg_.page_size);
      drv_->set_access(v_ptr_[layer] + first_off, done * cfg_.page_size);
    }
    if (pages_counter) *pages_counter += done * 2 * nl;
  }
}

void KvAllocator::unmap_one(u64 req) {
  // vattention.cu:219-241 + mux.h:51-66
  if (mapped_pages_[req] == 0) throw StateError("[vattn] unmap on a request with no pages");
  const u64 nl = cfg_.megacache ? 1 : cfg_.num_layers;
  const u64 off = req * cfg_.per_req + (mapped_pages_[req] - 1) * cfg_.page_size;  // utils.h:193-204
  if (fence_enabled_) drv_->wait_fence(fence_slot_);
  for (u64 layer = 0; layer < nl; layer++) {
    auto it = pagemap_.find(Key(req, off, layer));
    if (it == pagemap_.end()) throw StateError("[vattn] page map entry missing on unmap");
    if (logical()) {
      chunk_unref(k_ptr_[layer], off);
      chunk_unref(v_ptr_[layer], off);
    } else {
      drv_->unmap(k_ptr_[layer] + off, cfg_.page_size);
      drv_->unmap(v_ptr_[layer] + off, cfg_.page_size);
    }
    PhysPage pg[2] = {it->second.first, it->second.second};  // K pushed first, then V
    for (const PhysPage& p : pg) {
      auto sh = shared_refs_.find(p.id);
      if (sh != shared_refs_.end()) {
        if (--sh->second > 0) continue;  // still mapped under another request
        shared_refs_.erase(sh);
      }
      pool_.push_back(p);
    }
    pagemap_.erase(it);
  }
  mapped_pages_[req]--;
}

void KvAllocator::release_some(u64 req, u64 retain) {
  // vattention.cu:243-252
  while (mapped_pages_[req] > retain) unmap_one(req);
}

void KvAllocator::map_for_curr_step(u64 req, u64 seq_len, u64* pages_counter) {
  // vattention.cu:376-392
  u64 need = tokens_to_pages(seq_len);
  u64 have = mapped_pages_[req];
  if (need <= have) return;
  need -= have;
  if (!kvblocks_available(need)) reclaim_on_demand(need);
  log("[DEBUG] allocating " + std::to_string(need) + " pages for reqId: " + std::to_string(req));
  grow(req, need, true, pages_counter);
  seq_lens_[req] = seq_len;
}

void KvAllocator::reclaim_on_demand(u64 nblocks) {
  // vattention.cu:420-438: walk reqIds high -> low, free what is beyond need
  for (u64 i = cfg_.max_batch_size; i-- > 0;) {
    if (kvblocks_available(nblocks)) break;
    u64 have = mapped_pages_[i];
    u64 need = tokens_to_pages(seq_lens_[i]);
    if (have <= need) continue;
    release_some(i, need);
  }
}

void KvAllocator::do_reclaim_pages() {
  // vattention.cu:444-469
  if (deferred_reclaim_) return;
  int64_t next_prefill = -1;
  for (u64 r = 0; r < cfg_.max_batch_size; r++)
    if (seq_lens_[r] == 0) {
      next_prefill = static_cast<int64_t>(r);
      break;
    }
  for (u64 i = cfg_.max_batch_size; i-- > 0;) {
    if (seq_lens_[i] != 0 || static_cast<int64_t>(i) == next_prefill) continue;
    if (mapped_pages_[i] == 0) continue;
    unmap_one(i);
    break;
  }
}

u64 KvAllocator::need_new_page_async(u64 req, u64 eager) const {
  // utils.h:206-219
  if (seq_lens_[req] == 0) return 0;
  u64 have = mapped_pages_[req];
  if (have == cfg_.max_pages_per_req) return 0;
  u64 need = tokens_to_pages(seq_lens_[req] + eager);
  return need <= have ? 0 : need - have;
}

bool KvAllocator::pass_will_reclaim() const {
  // the first decision of background_pass, taken from the same state (nothing changes between the
  // hand-off and the pass)
  u64 nr_required = 0;
  for (u64 r = 0; r < cfg_.max_batch_size; r++) nr_required += need_new_page_async(r, 1);
  return !kvblocks_available(nr_required);
}

void KvAllocator::background_pass() {
  // vattention.cu:488-536
  u64 nr_required = 0;
  u64 nr_mapped_curr = 0;
  bool done = false;
  for (u64 r = 0; r < cfg_.max_batch_size; r++) nr_required += need_new_page_async(r, 1);
  if (!kvblocks_available(nr_required)) {
    log("[DEBUG] reclaiming " + std::to_string(nr_required) + " KV blocks in background thread...");
    reclaim_on_demand(nr_required);
  }
  if (!kvblocks_available(nr_required)) return;
  for (u64 eager = 1; eager < kEagerNumSteps && !done; eager++) {
    for (u64 r = 0; r < cfg_.max_batch_size; r++) {
      u64 n = need_new_page_async(r, eager);
      grow(r, n, false, &stats_.async_pages_mapped);
      nr_mapped_curr += n;
      if (eager == 1) continue;
      if (nr_mapped_curr >= kEagerNumKvBlocks) {
        done = true;
        break;
      }
    }
  }
  if (nr_required) return;
  do_reclaim_pages();
}

// ------------------------------------------------------------ mapper thread -

void KvAllocator::mapper_main() {
  bool bound = false;
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    cv_.wait(lk, [&] { return job_pending_ || stop_; });
    if (stop_) return;
    job_pending_ = false;
    // busy_ was set by step_async under the lock, so no API call can slip in between the hand-off
    // and this pass (reference hazard (a)).  While busy_ is set this thread owns the
    // bookkeeping; mu_ is released so that step_async can look at the hand-off snapshot and queue
    // the next step instead of standing still for the whole pass.
    for (;;) {
      lk.unlock();
      const u64 t0 = now_ns();
      std::string err;
      try {
        if (!bound) {
          drv_->bind_thread();
          bound = true;
        }
        stats_.async_pages_mapped = 0;
        background_pass();
      } catch (const std::exception& e) {
        err = e.what();
      }
      const u64 dt = now_ns() - t0;
      lk.lock();
      stats_.background_ns = dt;
      stats_.total_background_ns += dt;
      stats_.total_async_pages += stats_.async_pages_mapped;
      if (dt > stats_.max_background_ns) stats_.max_background_ns = dt;
      stats_.passes++;
      if (!err.empty()) bg_error_ = err;
      if (!queued_ || !bg_error_.empty()) break;
      // the queued step: what step_async does after its wait -- the lengths are replaced, preparing
      // the step maps nothing (that is what let it queue), the next pass starts
      queued_ = false;
      seq_lens_ = queued_lens_;
      fence_slot_ ^= 1;  // the queued step recorded its fence in the other slot
      handoff_lens_ = seq_lens_;
      handoff_mapped_ = mapped_pages_;
      handoff_reclaims_ = pass_will_reclaim();
      cv_.notify_all();  // a step_async waiting for the queue slot
    }
    queued_ = false;  // (only after an error: the queued lengths are dropped with it)
    busy_ = false;
    cv_.notify_all();
  }
}

// ------------------------------------------------------------------- steps --

void KvAllocator::step_sync(const u64* seq_lens, size_t n, bool eager_reclaim) {
  u64 t0 = now_ns();
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  if (n != cfg_.max_batch_size) throw InvalidError("[vattn] seq_lens must have max_batch_size entries");
  if (fence_enabled_) drv_->record_fence(compute_stream_, fence_slot_);
  stats_.sync_pages_mapped = 0;
  // vattention.cu:395-409
  for (u64 r = 0; r < cfg_.max_batch_size; r++) {
    seq_lens_[r] = seq_lens[r];
    if (eager_reclaim && seq_lens[r] == 0 && mapped_pages_[r] != 0) {
      release_some(r, 0);
      continue;
    }
    map_for_curr_step(r, seq_lens[r], &stats_.sync_pages_mapped);
  }
  stats_.critical_path_ns = now_ns() - t0;
  stats_.total_critical_path_ns += stats_.critical_path_ns;
  stats_.total_sync_pages += stats_.sync_pages_mapped;
  stats_.steps++;
}

void KvAllocator::hand_off_locked() {
  // spawn_kvcache_manager, vattention.cu:538-546 -> wake the parked mapper
  handoff_lens_ = seq_lens_;
  handoff_mapped_ = mapped_pages_;
  handoff_reclaims_ = pass_will_reclaim();
  busy_ = true;
  job_pending_ = true;
}

bool KvAllocator::can_ride_behind(const u64* seq_lens, size_t n) const {
  // The pass in flight started from handoff_lens_ / handoff_mapped_.  Unless it takes pages back
  // (handoff_reclaims_), it only ADDS pages to requests that were active at the hand-off and only
  // removes pages from requests that were not.  So if every request of the new step was active then
  // and already had the pages the new length needs, preparing this step would map nothing whenever
  // it ran: it can be queued behind the pass.
  if (handoff_reclaims_) return false;
  for (size_t r = 0; r < n; r++) {
    if (seq_lens[r] == 0) continue;
    if (handoff_lens_[r] == 0) return false;
    if (tokens_to_pages(seq_lens[r]) > handoff_mapped_[r]) return false;
  }
  return true;
}

void KvAllocator::step_async(const u64* seq_lens, size_t n) {
  u64 t0 = now_ns();
  std::unique_lock<std::mutex> lk(mu_);
  if (busy_ && queueing_ && configured_ && n == cfg_.max_batch_size && bg_error_.empty()) {
    // at most one step rides behind the pass in flight
    cv_.wait(lk, [&] { return !busy_ || !queued_; });
    if (busy_ && bg_error_.empty() && can_ride_behind(seq_lens, n)) {
      if (fence_enabled_) drv_->record_fence(compute_stream_, fence_slot_ ^ 1);  // may throw: nothing queued yet
      queued_lens_.assign(seq_lens, seq_lens + n);
      queued_ = true;
      stats_.sync_pages_mapped = 0;
      stats_.critical_path_ns = now_ns() - t0;
      stats_.total_critical_path_ns += stats_.critical_path_ns;
      stats_.steps++;
      stats_.queued_steps++;
      return;
    }
  }
  // vattention.cu:549-558, with the wait moved BEFORE the lengths are replaced
  // (reference hazard (b)): the previous pass must not see the new lengths.
  wait_idle(lk);
  require_configured();
  if (n != cfg_.max_batch_size) throw InvalidError("[vattn] seq_lens must have max_batch_size entries");
  if (fence_enabled_) drv_->record_fence(compute_stream_, fence_slot_);
  seq_lens_.assign(seq_lens, seq_lens + n);
  stats_.sync_pages_mapped = 0;
  // prepare_prefill_kvcache, vattention.cu:412-418
  for (u64 r = 0; r < cfg_.max_batch_size; r++) map_for_curr_step(r, seq_lens_[r], &stats_.sync_pages_mapped);
  hand_off_locked();
  stats_.critical_path_ns = now_ns() - t0;
  stats_.total_critical_path_ns += stats_.critical_path_ns;
  stats_.total_sync_pages += stats_.sync_pages_mapped;
  stats_.steps++;
  lk.unlock();
  cv_.notify_all();
}

int KvAllocator::alloc_new_batch_idx(u64 seqlen) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  // vattention.cu:564-589: best fit over inactive reqIds
  int new_id = -1;
  u64 need = tokens_to_pages(seqlen);
  for (u64 r = 0; r < cfg_.max_batch_size; r++) {
    if (seq_lens_[r] != 0) continue;
    if (new_id == -1) {
      new_id = static_cast<int>(r);
      continue;
    }
    if (mapped_pages_[r] >= need && mapped_pages_[r] < mapped_pages_[new_id]) new_id = static_cast<int>(r);
  }
  if (new_id != -1) seq_lens_[new_id] = seqlen;
  return new_id;
}

void KvAllocator::free_batch_idx(int req_id) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  if (req_id < 0 || static_cast<u64>(req_id) >= cfg_.max_batch_size)
    throw InvalidError("[vattn] reqId out of range");
  seq_lens_[req_id] = 0;  // vattention.cu:591-594: pages stay mapped (deferred reclamation)
}

u64 KvAllocator::num_free_kvblocks() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  return blocks_in_pool() + overcommitted();  // vattention.cu:194-211
}

void KvAllocator::map_common_pages(u64 num_tokens) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  // vattention.cu:326-373 + mux.h:68-85
  if (logical()) throw InvalidError("[vattn] map_common_pages is not available with logical (sub-granularity) pages");
  u64 nblocks = tokens_to_pages(num_tokens);
  if (nblocks == 0) return;
  if (!kvblocks_available(nblocks)) {
    log("free pages: " + std::to_string(blocks_in_pool()));
    log("required: " + std::to_string(nblocks));
    throw OomError(kOomMsg);
  }
  const u64 nl = cfg_.megacache ? 1 : cfg_.num_layers;
  for (u64 r = 0; r < cfg_.max_batch_size; r++)
    if (mapped_pages_[r] + nblocks > cfg_.max_pages_per_req)
      throw InvalidError("[vattn] map_common_pages would exceed a request's virtual range");
  for (u64 count = 0; count < nblocks; count++) {
    for (u64 layer = 0; layer < nl; layer++) {
      PhysPage k = pop_page();
      PhysPage v = pop_page();
      for (u64 r = 0; r < cfg_.max_batch_size; r++) {
        u64 off = r * cfg_.per_req + mapped_pages_[r] * cfg_.page_size;
        map_pair(r, layer, off, k, v);
        drv_->set_access(k_ptr_[layer] + off, cfg_.page_size);
        drv_->set_access(v_ptr_[layer] + off, cfg_.page_size);
      }
      // the reference pushes a shared handle back once per request on unmap
      // (mux.h:57-58), duplicating it in the pool; count the aliases instead.
      shared_refs_[k.id] = cfg_.max_batch_size;
      shared_refs_[v.id] = cfg_.max_batch_size;
    }
    for (u64 r = 0; r < cfg_.max_batch_size; r++) mapped_pages_[r]++;
  }
}

void KvAllocator::cleanup() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  if (!configured_) return;
  // vattention.cu:602-609, mux.h:24-35, cudaInternal.h:84-94
  for (u64 r = 0; r < cfg_.max_batch_size; r++) release_some(r, 0);
  for (size_t i = 0; i < k_ptr_.size(); i++) {
    drv_->addr_free(k_ptr_[i], cfg_.virt_size);
    drv_->addr_free(v_ptr_[i], cfg_.virt_size);
  }
  for (const PhysPage& p : pool_)
    if (!logical()) drv_->release(p.handle);
  for (u64 h : chunk_pool_) drv_->release(h);
  chunk_pool_.clear();
  chunks_.clear();
  slack_chunks_ = 0;
  pool_.clear();
  pagemap_.clear();
  shared_refs_.clear();
  k_ptr_.clear();
  v_ptr_.clear();
  mapped_pages_.clear();
  seq_lens_.clear();
  created_ = 0;
  configured_ = false;
  log("released memory and cleaned up vattention ...");
}

// ----------------------------------------------------------------- misc -----

void KvAllocator::set_verbose(bool v) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  verbose_ = v;
}

void KvAllocator::set_deferred_reclamation(bool v) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  deferred_reclaim_ = v;
}

void KvAllocator::show_kvcache_config() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  // vattention.cu:130-140
  log("Num layers: " + std::to_string(cfg_.num_layers));
  log("Num kv_heads: " + std::to_string(cfg_.num_kv_heads));
  log("Head size: " + std::to_string(cfg_.head_size));
  log("Max batch size: " + std::to_string(cfg_.max_batch_size));
  log("Max context length: " + std::to_string(cfg_.max_context_length));
  log("Bytes per elem: " + std::to_string(cfg_.bytes_per_elem));
  log("virt_buff_size_per_req: " + std::to_string(cfg_.per_req));
  log("virt_buff_size: " + std::to_string(cfg_.virt_size));
}

void KvAllocator::dump_state_locked() {
  // vattention.cu:76-95
  log("Free pool: " + std::to_string(blocks_in_pool()) + " KV blocks");
  log("reqId : seqlen: mapped: required");
  for (u64 i = 0; i < cfg_.max_batch_size; i++) {
    std::stringstream ss;
    ss << std::setw(8) << i << ": " << std::setw(8) << seq_lens_[i] << " : " << std::setw(8)
       << mapped_pages_[i] << " : " << std::setw(8) << tokens_to_pages(seq_lens_[i]);
    log(ss.str());
  }
}

void KvAllocator::show_allocator_state() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  if (!configured_) return;
  dump_state_locked();
}

void KvAllocator::wait_background() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
}

void KvAllocator::set_compute_stream(void* stream, bool enable) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  compute_stream_ = stream;
  fence_enabled_ = enable;
}

void KvAllocator::set_queueing(bool on) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  queueing_ = on;
}

StepStats KvAllocator::stats() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  return stats_;
}

KvConfig KvAllocator::config() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  return cfg_;
}

void KvAllocator::get_state(u64* mapped, u64* lens, size_t n) {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  require_configured();
  if (n != cfg_.max_batch_size) throw InvalidError("[vattn] state arrays must have max_batch_size entries");
  for (size_t i = 0; i < n; i++) {
    if (mapped) mapped[i] = mapped_pages_[i];
    if (lens) lens[i] = seq_lens_[i];
  }
}

std::vector<u64> KvAllocator::free_pool_ids() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  std::vector<u64> ids;
  ids.reserve(pool_.size());
  for (const PhysPage& p : pool_) ids.push_back(p.id);
  return ids;
}

std::vector<u64> KvAllocator::pagemap_words() {
  std::unique_lock<std::mutex> lk(mu_);
  wait_idle(lk);
  std::vector<u64> w;
  w.reserve(pagemap_.size() * 5);
  for (const auto& kv : pagemap_) {
    w.push_back(std::get<0>(kv.first));
    w.push_back(std::get<1>(kv.first));
    w.push_back(std::get<2>(kv.first));
    w.push_back(kv.second.first.id);
    w.push_back(kv.second.second.id);
  }
  return w;
}

}  // namespace vattn
