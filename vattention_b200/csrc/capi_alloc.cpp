// C ABI, part A (allocator).  Declarations and the reference lines each entry
// replaces are in include/vattn_b200.h.
#include <cstring>
#include <string>

#include "../../include/vattn_b200.h"
#include "capi_common.h"
#include "kv_allocator.h"

using vattn::KvAllocator;
using vattn::u64;

struct vattn_allocator {
  std::unique_ptr<KvAllocator> impl;
  vattn::MockVmmDriver* mock = nullptr;  // non-owning, set for HOST_MOCK
};

namespace vattn {
thread_local std::string g_last_error;

int translate_exception() {
  try {
    throw;
  } catch (const OomError& e) {
    g_last_error = e.what();
    return VATTN_ERR_OOM;
  } catch (const InvalidError& e) {
    g_last_error = e.what();
    return VATTN_ERR_INVALID;
  } catch (const StateError& e) {
    g_last_error = e.what();
    return VATTN_ERR_STATE;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return VATTN_ERR_DRIVER;
  } catch (...) {
    g_last_error = "unknown C++ exception";
    return VATTN_ERR_DRIVER;
  }
}
}  // namespace vattn

#define VATTN_TRY try {
#define VATTN_CATCH                      \
  }                                      \
  catch (...) {                          \
    return vattn::translate_exception(); \
  }

static bool check_handle(vattn_allocator_t* a) {
  if (a && a->impl) return true;
  vattn::g_last_error = "[vattn] null allocator handle";
  return false;
}

extern "C" {

const char* vattn_last_error(void) { return vattn::g_last_error.c_str(); }
const char* vattn_version(void) { return "vattn_b200 0.1 (sm_100a)"; }

int vattn_create(vattn_allocator_t** out, int backend) {
  if (!out) return VATTN_ERR_INVALID;
  VATTN_TRY
  auto* a = new vattn_allocator();
  if (backend == VATTN_BACKEND_HOST_MOCK) {
    auto m = std::make_unique<vattn::MockVmmDriver>();
    a->mock = m.get();
    a->impl = std::make_unique<KvAllocator>(std::move(m));
  } else if (backend == VATTN_BACKEND_CUDA) {
    std::unique_ptr<vattn::VmmDriver> d;
    try {
      d = vattn::make_cuda_vmm_driver();
    } catch (...) {
      delete a;
      throw;
    }
    a->impl = std::make_unique<KvAllocator>(std::move(d));
  } else {
    delete a;
    vattn::g_last_error = "[vattn] unknown backend";
    return VATTN_ERR_INVALID;
  }
  *out = a;
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_destroy(vattn_allocator_t* a) {
  if (!a) return VATTN_OK;
  // the handle is freed even when cleanup() rethrows a stored mapper-thread error
  struct Reaper {
    vattn_allocator_t* h;
    ~Reaper() { delete h; }
  } reaper{a};
  VATTN_TRY
  if (a->impl) a->impl->cleanup();
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_init_kvcache(vattn_allocator_t* a, uint64_t num_layers, uint64_t num_kv_heads,
                       uint64_t head_size, uint64_t max_batch_size, uint64_t max_context_length,
                       int device, uint64_t bytes_per_elem, uint64_t page_size, int megacache,
                       uint64_t* ptrs, int* n_ptrs, int64_t shape[5], int* ndim) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  std::vector<u64> p = a->impl->init_kvcache(num_layers, num_kv_heads, head_size, max_batch_size,
                                             max_context_length, device, bytes_per_elem, page_size,
                                             megacache != 0);
  if (ptrs)
    for (size_t i = 0; i < p.size(); i++) ptrs[i] = p[i];
  if (n_ptrs) *n_ptrs = static_cast<int>(p.size());
  if (shape && ndim) {
    // vattention.cu:145-149
    if (megacache) {
      int64_t s[5] = {(int64_t)max_batch_size, (int64_t)max_context_length, (int64_t)num_layers,
                      (int64_t)num_kv_heads, (int64_t)head_size};
      std::memcpy(shape, s, sizeof(s));
      *ndim = 5;
    } else {
      int64_t s[5] = {(int64_t)max_batch_size, (int64_t)max_context_length, (int64_t)num_kv_heads,
                      (int64_t)head_size, 0};
      std::memcpy(shape, s, sizeof(s));
      *ndim = 4;
    }
  }
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_get_config(vattn_allocator_t* a, vattn_config_t* out) {
  if (!check_handle(a) || !out) return VATTN_ERR_INVALID;
  VATTN_TRY
  vattn::KvConfig c = a->impl->config();
  out->num_layers = c.num_layers;
  out->num_kv_heads = c.num_kv_heads;
  out->head_size = c.head_size;
  out->max_batch_size = c.max_batch_size;
  out->max_context_length = c.max_context_length;
  out->bytes_per_elem = c.bytes_per_elem;
  out->page_size = c.page_size;
  out->megacache = c.megacache;
  out->tokens_per_page = c.tokens_per_page;
  out->virt_buff_size_per_token = c.per_token;
  out->virt_buff_size_per_req = c.per_req;
  out->virt_buff_size = c.virt_size;
  out->max_pages_per_req = c.max_pages_per_req;
  out->phys_granularity = c.granularity;
  out->num_tensors = c.megacache ? 2 : 2 * c.num_layers;
  return VATTN_OK;
  VATTN_CATCH
}

int64_t vattn_reserve_physical_pages(vattn_allocator_t* a, uint64_t free_memory) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  return static_cast<int64_t>(a->impl->reserve_physical_pages(free_memory));
  VATTN_CATCH
}

int vattn_step(vattn_allocator_t* a, const uint64_t* seq_lens, size_t n, int eager_reclaim) {
  if (!check_handle(a) || !seq_lens) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->step_sync(seq_lens, n, eager_reclaim != 0);
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_step_async(vattn_allocator_t* a, const uint64_t* seq_lens, size_t n) {
  if (!check_handle(a) || !seq_lens) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->step_async(seq_lens, n);
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_alloc_new_batch_idx(vattn_allocator_t* a, uint64_t seqlen) {
  if (!check_handle(a)) return VATTN_ERR_INVALID - 100;
  try {
    return a->impl->alloc_new_batch_idx(seqlen);
  } catch (...) {
    // -1 is the reference's "no slot" answer (vattention.cu:567); errors are < -1
    return vattn::translate_exception() - 100;
  }
}

int vattn_free_batch_idx(vattn_allocator_t* a, int req_id) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->free_batch_idx(req_id);
  return VATTN_OK;
  VATTN_CATCH
}

uint64_t vattn_num_free_kvblocks(vattn_allocator_t* a) {
  if (!check_handle(a)) return 0;
  try {
    return a->impl->num_free_kvblocks();
  } catch (...) {
    vattn::translate_exception();
    return 0;
  }
}

int vattn_cleanup(vattn_allocator_t* a) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->cleanup();
  return VATTN_OK;
  VATTN_CATCH
}

void vattn_set_verbose(vattn_allocator_t* a, int on) {
  if (check_handle(a)) try {
      a->impl->set_verbose(on != 0);
    } catch (...) {
      vattn::translate_exception();
    }
}

void vattn_set_deferred_reclamation(vattn_allocator_t* a, int on) {
  if (check_handle(a)) try {
      a->impl->set_deferred_reclamation(on != 0);
    } catch (...) {
      vattn::translate_exception();
    }
}

void vattn_show_kvcache_config(vattn_allocator_t* a) {
  if (check_handle(a)) try {
      a->impl->show_kvcache_config();
    } catch (...) {
      vattn::translate_exception();
    }
}

void vattn_show_allocator_state(vattn_allocator_t* a) {
  if (check_handle(a)) try {
      a->impl->show_allocator_state();
    } catch (...) {
      vattn::translate_exception();
    }
}

int vattn_map_common_pages(vattn_allocator_t* a, uint64_t num_tokens) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->map_common_pages(num_tokens);
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_wait_background(vattn_allocator_t* a) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->wait_background();
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_set_compute_stream(vattn_allocator_t* a, void* stream, int enable) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->set_compute_stream(stream, enable != 0);
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_get_step_stats(vattn_allocator_t* a, vattn_step_stats_t* out) {
  if (!check_handle(a) || !out) return VATTN_ERR_INVALID;
  VATTN_TRY
  vattn::StepStats s = a->impl->stats();
  out->critical_path_ns = s.critical_path_ns;
  out->background_ns = s.background_ns;
  out->sync_pages_mapped = s.sync_pages_mapped;
  out->async_pages_mapped = s.async_pages_mapped;
  out->driver_calls = a->impl->driver()->calls();
  out->total_critical_path_ns = s.total_critical_path_ns;
  out->total_background_ns = s.total_background_ns;
  out->max_background_ns = s.max_background_ns;
  out->total_sync_pages = s.total_sync_pages;
  out->total_async_pages = s.total_async_pages;
  out->steps = s.steps;
  out->passes = s.passes;
  out->queued_steps = s.queued_steps;
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_set_queueing(vattn_allocator_t* a, int on) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->set_queueing(on != 0);
  return VATTN_OK;
  VATTN_CATCH
}

int vattn_get_state(vattn_allocator_t* a, uint64_t* mapped_pages, uint64_t* seq_lens, size_t n) {
  if (!check_handle(a)) return VATTN_ERR_INVALID;
  VATTN_TRY
  a->impl->get_state(mapped_pages, seq_lens, n);
  return VATTN_OK;
  VATTN_CATCH
}

size_t vattn_get_free_pool(vattn_allocator_t* a, uint64_t* ids, size_t cap) {
  if (!check_handle(a)) return 0;
  try {
    std::vector<u64> v = a->impl->free_pool_ids();
    for (size_t i = 0; i < v.size() && i < cap && ids; i++) ids[i] = v[i];
    return v.size();
  } catch (...) {
    vattn::translate_exception();
    return 0;
  }
}

size_t vattn_get_pagemap(vattn_allocator_t* a, uint64_t* words, size_t cap_entries) {
  if (!check_handle(a)) return 0;
  try {
    std::vector<u64> w = a->impl->pagemap_words();
    size_t n = w.size() / 5;
    for (size_t i = 0; i < n && i < cap_entries && words; i++)
      std::memcpy(words + 5 * i, w.data() + 5 * i, 5 * sizeof(u64));
    return n;
  } catch (...) {
    vattn::translate_exception();
    return 0;
  }
}

size_t vattn_get_driver_log(vattn_allocator_t* a, uint64_t* words, size_t cap_records) {
  if (!check_handle(a) || !a->mock) return 0;
  a->impl->wait_background();
  std::vector<vattn::DriverLogRecord> log = a->mock->snapshot_log();
  for (size_t i = 0; i < log.size() && i < cap_records && words; i++) {
    words[4 * i + 0] = log[i].op;
    words[4 * i + 1] = log[i].va;
    words[4 * i + 2] = log[i].size;
    words[4 * i + 3] = log[i].handle;
  }
  return log.size();
}

void vattn_mock_set_capacity(vattn_allocator_t* a, uint64_t bytes) {
  if (check_handle(a) && a->mock) {
    a->impl->wait_background();
    a->mock->set_capacity(bytes);
  }
}

void vattn_mock_set_call_delay_us(vattn_allocator_t* a, uint64_t us) {
  if (check_handle(a) && a->mock) a->mock->set_call_delay_us(us);
}

void vattn_mock_fence_counts(vattn_allocator_t* a, uint64_t out[4]) {
  if (!check_handle(a) || !a->mock || !out) return;
  a->impl->wait_background();
  out[0] = a->mock->fence_records(0), out[1] = a->mock->fence_records(1);
  out[2] = a->mock->fence_waits(0), out[3] = a->mock->fence_waits(1);
}

void vattn_clear_driver_log(vattn_allocator_t* a) {
  if (check_handle(a) && a->mock) {
    a->impl->wait_background();
    a->mock->clear_log();
  }
}

}  // extern "C"
