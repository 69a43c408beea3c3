// ThreadSanitizer harness for the allocator's two threads (API thread + mapper), host mock driver.
// Built and run by tests/test_allocator_tsan.py:
//   g++ -fsanitize=thread -O1 -g -std=c++17 allocator_tsan.cpp ../../vattention_b200/csrc/{kv_allocator,vmm_driver,capi_alloc}.cpp
// Drives decode traces through the C ABI with a slow mock driver so that step_async calls are queued
// behind passes in flight, interleaved with the calls that wait for the mapper.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/vattn_b200.h"

#define CHECK(x)                                                        \
  do {                                                                  \
    if ((x) < 0) {                                                      \
      std::fprintf(stderr, "FAILED %s: %s\n", #x, vattn_last_error()); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main() {
  vattn_allocator_t* a = nullptr;
  CHECK(vattn_create(&a, VATTN_BACKEND_HOST_MOCK));
  const uint64_t L = 2, Hkv = 2, D = 64, B = 6, ctx = 32768, page = 2ull << 20;
  uint64_t ptrs[2 * L];
  int n_ptrs = 0, ndim = 0;
  int64_t shape[5];
  CHECK(vattn_init_kvcache(a, L, Hkv, D, B, ctx, 0, 2, page, 0, ptrs, &n_ptrs, shape, &ndim));
  CHECK((int)vattn_reserve_physical_pages(a, 160 * page));
  CHECK(vattn_set_compute_stream(a, nullptr, 1));
  const uint64_t tpp = page / (Hkv * D * 2);
  std::vector<uint64_t> lens(B, 0);
  unsigned rng = 12345;
  auto rnd = [&] { return rng = rng * 1664525u + 1013904223u, rng >> 8; };
  vattn_mock_set_call_delay_us(a, 100);
  uint64_t queued = 0;
  for (int it = 0; it < 400; it++) {
    const unsigned r = rnd() % 100;
    if (r < 20) {
      const uint64_t n = (rnd() % 3 == 0) ? tpp - (rnd() % 12) : 1 + rnd() % (2 * tpp);
      const int id = vattn_alloc_new_batch_idx(a, n);
      if (id >= 0) lens[id] = n;
    } else if (r < 28) {
      const int id = (int)(rnd() % B);
      if (lens[id]) {
        CHECK(vattn_free_batch_idx(a, id));
        lens[id] = 0;
      }
    } else if (r < 32) {
      vattn_step_stats_t st;
      CHECK(vattn_get_step_stats(a, &st));
      queued = st.queued_steps;
    }
    const int steps = 1 + (int)(rnd() % 6);
    for (int s = 0; s < steps; s++) {
      for (uint64_t i = 0; i < B; i++)
        if (lens[i] && lens[i] < ctx - 1) lens[i]++;
      CHECK(vattn_step_async(a, lens.data(), B));
    }
  }
  vattn_step_stats_t st;
  CHECK(vattn_get_step_stats(a, &st));
  std::printf("steps %llu passes %llu queued %llu async pages %llu\n", (unsigned long long)st.steps,
              (unsigned long long)st.passes, (unsigned long long)st.queued_steps,
              (unsigned long long)st.total_async_pages);
  CHECK(vattn_cleanup(a));
  CHECK(vattn_destroy(a));
  return st.queued_steps > 0 ? 0 : 2;
}
