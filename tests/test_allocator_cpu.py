"""Allocator bookkeeping, bit-exact against the oracle (oracle/allocator_model.py), on CPU.

The C++ allocator runs against the HOST_MOCK VMM driver (no GPU); the same API trace is
replayed on the oracle; mapped_pages, seq_lens, free-pool order, page map, return values
and num_free_kvblocks must be identical after every call.  The driver-call log is checked
too: cuMemMap / cuMemUnmap / cuMemCreate happen in the reference's order with the
reference's (tensor, offset, page) triples; cuMemSetAccess covers exactly what was mapped.
"""
import random

import pytest
import torch

from oracle.allocator_model import MB, AllocatorModel, AllocatorOOM, page_size_to_block_tokens
from vattention_b200 import _lib
from vattention_b200 import vattention as va


@pytest.fixture()
def mock_backend():
    va._use_backend(_lib.BACKEND_HOST_MOCK)
    yield
    va.cleanup()
    va._use_backend(_lib.BACKEND_CUDA)


def make_pair(L, Hkv, D, B, ctx, mega=False, mem_pages=None, itemsize=2, page=2 * MB):
    model = AllocatorModel(L, Hkv, D, B, ctx, itemsize, page, mega)
    dtype = {2: torch.bfloat16, 4: torch.float32}[itemsize]
    tensors = va.init_kvcache(L, Hkv, D, B, ctx, 0, dtype, page, mega)
    if mem_pages is not None:
        got = va.reserve_physical_pages(mem_pages * page)
        want = model.reserve_physical_pages(mem_pages * page)
        assert got == want
    return model, tensors


def assert_same(model):
    got = va.get_state()
    want = model.snapshot()
    assert got["mapped_pages"] == want["mapped_pages"]
    assert got["seq_lens"] == want["seq_lens"]
    assert got["pool"] == want["pool"]
    assert got["pagemap"] == want["pagemap"]
    assert got["num_free_kvblocks"] == want["num_free_kvblocks"]


def test_config_arithmetic(mock_backend):
    # Llama-3-8B shapes, SURVEY 8(a2): 2048 B/token -> 1024 tok/page, 32 pages/req at 32K
    model, tensors = make_pair(32, 8, 128, 64, 32768)
    cfg = va.get_config()
    assert cfg["tokens_per_page"] == model.tokens_per_page == 1024
    assert cfg["virt_buff_size_per_token"] == 2048
    assert cfg["virt_buff_size_per_req"] == model.virt_buff_size_per_req == 64 * MB
    assert cfg["max_pages_per_req"] == model.max_pages_per_req == 32
    assert cfg["virt_buff_size"] == 4 * 1024 * MB
    assert len(tensors) == 64 and tuple(tensors[0].shape) == (64, 32768, 8, 128)
    # Yi-6B: 1024 B/token -> 2048 tok/page; 70B TP8: 256 B/token -> 8192 tok/page
    assert AllocatorModel(32, 4, 128, 8, 131072, 2).tokens_per_page == 2048
    assert AllocatorModel(80, 1, 128, 8, 131072, 2).tokens_per_page == 8192
    # engine-side conversion (arg_utils.py:147-159)
    assert page_size_to_block_tokens(2 * MB, 8, 128, 1, 32, False) == 1024
    assert page_size_to_block_tokens(2 * MB, 8, 128, 8, 80, False) == 8192
    assert page_size_to_block_tokens(2 * MB, 8, 128, 1, 32, True) == 32


def test_megacache_geometry(mock_backend):
    model, tensors = make_pair(4, 2, 64, 4, 16384, mega=True, mem_pages=64)
    assert len(tensors) == 2 and tuple(tensors[0].shape) == (4, 16384, 4, 2, 64)
    assert va.get_config()["tokens_per_page"] == model.tokens_per_page == 2 * MB // (2 * 64 * 2 * 4)
    assert_same(model)


def test_reserve_rounds_to_2L(mock_backend):
    model, _ = make_pair(3, 2, 64, 4, 8192)
    for mem in (0, 5 * MB, 12 * MB + 7, 13 * MB, 64 * MB):
        assert va.reserve_physical_pages(mem) == model.reserve_physical_pages(mem)
        assert_same(model)


def test_rejects_bad_config(mock_backend):
    with pytest.raises(RuntimeError, match="VMM granularity"):
        va.init_kvcache(2, 2, 64, 2, 8192, 0, torch.float16, 3 * MB, False)
    with pytest.raises(RuntimeError, match="must divide the device VMM granularity"):
        va.init_kvcache(2, 2, 64, 2, 8192, 0, torch.float16, 384 * 1024, False)
    with pytest.raises(RuntimeError, match="never spans two requests"):   # per_req = 1 MiB
        va.init_kvcache(2, 2, 64, 2, 4096, 0, torch.float16, 256 * 1024, False)
    with pytest.raises(RuntimeError, match="multiple of page_size"):
        va.init_kvcache(2, 2, 64, 2, 1000, 0, torch.float16, 2 * MB, False)
    with pytest.raises(RuntimeError, match="max_batch_size"):
        va.init_kvcache(2, 2, 64, 1000, 8192, 0, torch.float16, 2 * MB, False)
    with pytest.raises(RuntimeError, match="init_kvcache has not been called"):
        va.step([0, 0], True)


def test_sync_step_trace(mock_backend):
    model, _ = make_pair(2, 2, 64, 4, 32768, mem_pages=40)
    tpp = model.tokens_per_page
    lens = [0, 0, 0, 0]
    for lens in ([tpp, 0, 0, 0], [tpp + 1, 5, 0, 0], [3 * tpp, 5, 2 * tpp, 0], [3 * tpp, 0, 0, 7],
                 [0, 0, 0, 0], [1, 1, 1, 1]):
        va.step(lens, True)
        model.step(lens, True)
        assert_same(model)
    va.step([2 * tpp, 0, 0, 0], False)
    model.step([2 * tpp, 0, 0, 0], False)
    assert_same(model)


def test_oom_message_and_state(mock_backend):
    model, _ = make_pair(2, 2, 64, 2, 32768, mem_pages=8)  # 2 KV blocks
    tpp = model.tokens_per_page
    with pytest.raises(RuntimeError, match="OOM on demand: not enough free pages"):
        va.step([3 * tpp, 0], False)
    with pytest.raises(AllocatorOOM):
        model.step([3 * tpp, 0], False)
    assert_same(model)
    va.set_verbose(False)


def test_best_fit_reqid(mock_backend):
    model, _ = make_pair(1, 2, 64, 4, 65536, mem_pages=64)
    tpp = model.tokens_per_page
    for lens in ([4 * tpp, 2 * tpp, 1, 3 * tpp],):
        va.step(lens, False)
        model.step(lens, False)
    for r in range(4):
        va.free_batch_idx(r)
        model.free_batch_idx(r)
    # all inactive with 4,2,1,3 pages mapped: best fit picks the smallest sufficient
    for need in (2 * tpp, 3 * tpp, 1, 10 * tpp, 5):
        got, want = va.alloc_new_batch_idx(need), model.alloc_new_batch_idx(need)
        assert got == want
        assert_same(model)


def random_trace(seed, B, ctx, tpp, steps):
    rng = random.Random(seed)
    lens = [0] * B
    ops = []
    for _ in range(steps):
        r = rng.random()
        active = [i for i in range(B) if lens[i]]
        if r < 0.25 and len(active) < B:
            n = rng.choice([1, tpp - 1, tpp, tpp + 1, 3 * tpp + 17, rng.randrange(1, ctx // 2)])
            ops.append(("alloc", n))
        elif r < 0.35 and active:
            ops.append(("free", rng.choice(active)))
        elif r < 0.40:
            ops.append(("free_blocks",))
        else:
            ops.append(("step", rng.randrange(1, 4)))
    return ops


@pytest.mark.parametrize("mega", [False, True])
@pytest.mark.parametrize("mode", ["async", "sync", "async_nodefer"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_traces_match_oracle(mock_backend, mega, mode, seed):
    replay_random_trace(mega, mode, seed, 2 * MB, 12 if mega else 90)


def check_chunk_invariants():
    """Logical (sub-granularity) pages: every live logical page lies inside a mapped 2 MiB chunk,
    no chunk is mapped without a live page inside, all driver traffic is chunk sized."""
    cfg, st, log = va.get_config(), va.get_state(), va.get_driver_log()
    gran, page, per_req = 2 * MB, cfg["page_size"], cfg["virt_buff_size_per_req"]
    reserves = [r for r in log if r[0] == 1]
    nt = len(reserves) // 2
    live = set()
    for r in log:
        if r[0] == 3:
            assert r[2] == gran and r[1] % gran == 0 and r[1] not in live
            live.add(r[1])
        elif r[0] == 5:
            assert r[2] == gran
            live.remove(r[1])
    want = set()
    for req, off, layer, _, _ in st["pagemap"]:
        assert off // per_req == req
        for t in (reserves[layer][1], reserves[nt + layer][1]):
            want.add((t + off) // gran * gran)
    assert live == want
    granted = {r[1] for r in log if r[0] == 4}
    assert live <= granted


@pytest.mark.parametrize("page_kb", [64, 256, 1024])
@pytest.mark.parametrize("mode", ["async", "sync"])
def test_logical_pages_match_oracle(mock_backend, page_kb, mode):
    """64 KB ... 1 MB pages (the reference's UVM modes, utils.h:83-86): the bookkeeping is the
    reference's, physical memory moves in 2 MiB chunks."""
    page = page_kb * 1024
    replay_random_trace(False, mode, 3, page, 90 * (2 * MB // page), after_step=check_chunk_invariants)
    for i in range(6):
        va.free_batch_idx(i)
    va.step([0] * 6, True)
    check_chunk_invariants()
    assert not va.get_state()["pagemap"]


def test_logical_pages_megacache_and_common(mock_backend):
    page = 256 * 1024
    replay_random_trace(True, "async", 5, page, 12 * 8, after_step=check_chunk_invariants)
    with pytest.raises(RuntimeError, match="map_common_pages is not available"):
        va.map_common_pages(100)


def replay_random_trace(mega, mode, seed, page, mem_pages, after_step=None):
    L, Hkv, D, B, ctx = 3, 2, 64, 6, 32768
    model, _ = make_pair(L, Hkv, D, B, ctx, mega=mega, mem_pages=mem_pages, page=page)
    if mode == "async_nodefer":
        va.set_deferred_reclamation(False)
        model.set_deferred_reclamation(False)
    tpp = model.tokens_per_page
    lens = [0] * B
    rng = random.Random(100 + seed)
    for op in random_trace(seed, B, ctx, tpp, 160):
        if op[0] == "alloc":
            got, want = va.alloc_new_batch_idx(op[1]), model.alloc_new_batch_idx(op[1])
            assert got == want
            if got >= 0:
                lens[got] = op[1]
        elif op[0] == "free":
            va.free_batch_idx(op[1])
            model.free_batch_idx(op[1])
            lens[op[1]] = 0
        elif op[0] == "free_blocks":
            assert va.num_free_kvblocks() == model.num_free_kvblocks()
        else:
            for _ in range(op[1]):
                for i in range(B):           # decode: every active sequence grows by one token
                    if lens[i] and lens[i] < ctx - 1:
                        lens[i] += rng.choice([1, 1, 1, tpp // 2])
                        lens[i] = min(lens[i], ctx - 1)
                err_got = err_want = None
                try:
                    va.step_async(lens) if mode != "sync" else va.step(lens, True)
                except RuntimeError as e:
                    err_got = str(e)
                try:
                    model.step_async(lens) if mode != "sync" else model.step(lens, True)
                except AllocatorOOM as e:
                    err_want = str(e)
                assert (err_got is None) == (err_want is None)
                if err_want:
                    assert err_want in err_got
                    va.set_verbose(False)
                    # an OOM leaves both sides mid-step in the same state; drop the biggest
                    # request on both and carry on (what the scheduler's preemption does)
                    big = max(range(B), key=lambda i: lens[i])
                    lens[big] = 0
                    va.free_batch_idx(big)
                    model.free_batch_idx(big)
                assert_same(model)
                if after_step:
                    va.wait_background()
                    after_step()
        assert_same(model)


def test_driver_call_order_matches_reference(mock_backend):
    """map/unmap/create sequence == the reference's (oracle log); set_access covers the same
    bytes with fewer calls (one per contiguous range per tensor)."""
    L, B = 2, 3
    model, _ = make_pair(L, 2, 64, B, 32768, mem_pages=48)
    cfg = va.get_config()
    tpp, page, per_req = cfg["tokens_per_page"], cfg["page_size"], cfg["virt_buff_size_per_req"]
    log0 = va.get_driver_log()
    reserves = [r for r in log0 if r[0] == 1]
    assert len(reserves) == 2 * L and all(r[2] == per_req * B for r in reserves)
    base = {f"k{i}": reserves[i][1] for i in range(L)}
    base.update({f"v{i}": reserves[L + i][1] for i in range(L)})
    creates = [r for r in log0 if r[0] == 2]
    assert [r[3] for r in creates] == list(range(1, 49))  # mock handle ids are creation order + 1
    va.clear_driver_log()
    model.calls.clear()

    for lens in ([3 * tpp, 0, tpp], [3 * tpp + 1, 5, tpp], [0, 5, 0]):
        va.step(lens, True)
        model.step(lens, True)
    got = va.get_driver_log()
    want = model.calls
    got_maps = [(r[1], r[3] - 1) for r in got if r[0] == 3]
    want_maps = [(base[c.tensor] + c.offset, c.page) for c in want if c.op == "map"]
    assert got_maps == want_maps
    got_unmaps = [r[1] for r in got if r[0] == 5]
    want_unmaps = [base[c.tensor] + c.offset for c in want if c.op == "unmap"]
    assert got_unmaps == want_unmaps
    assert all(r[2] == page for r in got if r[0] in (3, 5))

    def covered(ranges):
        s = set()
        for va_, size in ranges:
            s.update(range(va_, va_ + size, page))
        return s
    got_acc = [(r[1], r[2]) for r in got if r[0] == 4]
    want_acc = [(base[c.tensor] + c.offset, page) for c in want if c.op == "set_access"]
    assert covered(got_acc) == covered(want_acc)
    assert len(got_acc) < len(want_acc)


def test_async_eager_mapping_policy(mock_backend):
    """The background pass maps what len+1 needs, then looks ahead up to +9 until two more
    blocks were requested (vattention.cu:513-526)."""
    model, _ = make_pair(1, 2, 64, 4, 32768, mem_pages=40)
    tpp = model.tokens_per_page
    lens = [tpp, tpp - 5, 3, 0]
    va.step_async(lens)
    model.step_async(lens)
    va.wait_background()
    assert_same(model)
    st = va.get_state()
    # req 0 is exactly at a page boundary -> the background pass already mapped its next page
    assert st["mapped_pages"][0] == 2
    # req 1 crosses within the 9-token look-ahead window -> mapped eagerly as well
    assert st["mapped_pages"][1] == 2
    assert st["mapped_pages"][2] == 1 and st["mapped_pages"][3] == 0
    stats = va.get_step_stats()
    assert stats["sync_pages_mapped"] == 2 * 3 and stats["async_pages_mapped"] == 2 * 2


def test_map_common_pages_refcounted(mock_backend):
    model, _ = make_pair(2, 2, 64, 3, 32768, mem_pages=24)
    tpp = model.tokens_per_page
    va.map_common_pages(tpp + 1)  # 2 blocks shared by all 3 requests
    st = va.get_state()
    assert st["mapped_pages"] == [2, 2, 2]
    assert len(st["pool"]) == 24 - 2 * 2 * 2       # 2 blocks x 2L pages, NOT x3 requests
    by_req = {}
    for req, off, layer, k, v in st["pagemap"]:
        by_req.setdefault((off % va.get_config()["virt_buff_size_per_req"], layer), set()).add((k, v))
    assert all(len(s) == 1 for s in by_req.values())  # same physical pair under every request
    va.step([0, 0, 0], True)                            # eager reclaim unmaps everything
    st = va.get_state()
    assert st["mapped_pages"] == [0, 0, 0]
    assert sorted(st["pool"]) == list(range(24))        # every page back exactly once


def test_cleanup_releases_everything(mock_backend):
    model, _ = make_pair(2, 2, 64, 2, 32768, mem_pages=16)
    va.step([model.tokens_per_page * 2, 3], False)
    va.clear_driver_log()
    va.cleanup()
    # re-init works after cleanup (the reference's globals cannot do this)
    va.init_kvcache(1, 1, 64, 1, 16384, 0, torch.float16, 2 * MB, False)
    log = va.get_driver_log()
    assert [r[0] for r in log].count(1) == 2


def test_logical_pages_cleanup_balances_the_driver(mock_backend):
    """256 KB logical pages: after cleanup every chunk that was created is released, every chunk
    that was mapped is unmapped, every reservation freed (driver log of the mock)."""
    page = 256 * 1024
    model, _ = make_pair(2, 2, 64, 4, 32768, mem_pages=64, page=page)
    tpp = model.tokens_per_page
    va.step([5 * tpp + 3, tpp, 0, 9 * tpp], True)
    va.step([5 * tpp + 3, 0, 17, 9 * tpp], True)
    va.cleanup()
    log = va.get_driver_log()
    n = lambda op: sum(1 for r in log if r[0] == op)
    assert n(2) > 0 and n(2) == n(6)          # create == release
    assert n(3) > 0 and n(3) == n(5)          # map == unmap
    assert n(1) == n(7) == 4                   # reserve == addr_free (2 layers x K, V)
    va.init_kvcache(1, 1, 64, 1, 16384, 0, torch.float16, 2 * MB, False)   # leave a live config for the fixture


def test_logical_reserve_stays_inside_the_budget(mock_backend):
    """with sub-granularity pages reserve_physical_pages must not create more physical
    memory than it was given; the per-(request, tensor) slack chunk is created only when a partly
    used chunk actually appears, and never more than max_batch_size x tensors of them."""
    page = 256 * 1024
    L, B = 4, 8
    budget = 64 * 2 * MB
    va.init_kvcache(L, 2, 64, B, 32768, 0, torch.bfloat16, page, False)
    va.clear_driver_log()
    n = va.reserve_physical_pages(budget)
    created = sum(r[2] for r in va.get_driver_log() if r[0] == 2)
    assert n == budget // page - (budget // page) % (2 * L)
    assert created <= budget and created >= n * page
    # one token in every request: each (request, tensor) takes a whole chunk for one logical page
    va.clear_driver_log()
    va.step([1] * B, True)
    extra = sum(r[2] for r in va.get_driver_log() if r[0] == 2)
    assert extra == 0                      # 64 chunks budgeted, 8 requests x 8 tensors = 64 needed
    tpp = va.get_config()["tokens_per_page"]
    # fill until the budgeted chunks are gone: further partly used chunks come from the slack
    lens = [1] * B
    with pytest.raises(RuntimeError):
        for _ in range(200):
            lens = [x + tpp for x in lens]
            va.step(lens, True)
    total = sum(r[2] for r in va.get_driver_log() if r[0] == 2)
    assert total <= B * 2 * L * 2 * MB     # never more than one slack chunk per (request, tensor)
    st = va.get_state()
    # after the OOM the books still balance: every logical page is either free or in the page map
    in_map = sum(2 for _ in st["pagemap"])
    assert in_map + len(st["pool"]) == n


def test_reserve_failure_publishes_nothing_and_can_be_retried(mock_backend):
    page = 256 * 1024
    va.init_kvcache(2, 2, 64, 4, 32768, 0, torch.bfloat16, page, False)
    va.mock_set_capacity(10 * 2 * MB)
    with pytest.raises(RuntimeError, match="cuMemCreate failed"):
        va.reserve_physical_pages(32 * 2 * MB)
    assert va.get_state()["pool"] == []               # no logical page without a chunk behind it
    assert va.num_free_kvblocks() == 0
    va.mock_set_capacity(0)
    assert va.reserve_physical_pages(32 * 2 * MB) == 32 * 8
    va.step([5, 0, 0, 0], True)
    assert va.get_state()["mapped_pages"] == [1, 0, 0, 0]


def test_grow_rolls_back_when_the_driver_fails_mid_block(mock_backend):
    """2 MiB pages: cuMemCreate is not on the grow path, so fail the slack chunk creation of the
    logical mode instead -- block k of a request fails in layer 1 after layer 0 was mapped."""
    page = 256 * 1024
    L, B = 2, 2
    va.init_kvcache(L, 2, 64, B, 32768, 0, torch.bfloat16, page, False)
    n = va.reserve_physical_pages(4 * 2 * MB)        # 4 chunks = exactly one per (request 0, tensor)
    va.step([1, 0], True)                            # request 0 holds all 4 chunks, one page each
    before = va.get_state()
    # request 1 now needs slack chunks; let the device hold only one more -> K of layer 0 maps,
    # V of layer 0 fails
    va.mock_set_capacity(5 * 2 * MB)
    with pytest.raises(RuntimeError, match="cuMemCreate failed"):
        va.step([1, 1], True)
    after = va.get_state()
    assert after["mapped_pages"] == before["mapped_pages"]
    assert after["pool"] == before["pool"]           # LIFO order restored
    assert after["pagemap"] == before["pagemap"]
    va.mock_set_capacity(0)
    va.step([1, 1], True)                            # and the same step succeeds afterwards
    assert va.get_state()["mapped_pages"] == [1, 1]
    assert len(va.get_state()["pool"]) == n - 2 * 2 * L


# ---- step_async queued behind a mapper pass that is still in flight ------------------------------

def _decode_steps(model, lens, steps, ctx, read_state_every=0):
    """`steps` decode iterations back to back: no call that waits for the mapper in between."""
    import time
    took = []
    for s in range(steps):
        lens = [min(n + 1, ctx - 1) if n else 0 for n in lens]
        t0 = time.perf_counter()
        va.step_async(lens)
        took.append(time.perf_counter() - t0)
        model.step_async(lens)
        if read_state_every and (s + 1) % read_state_every == 0:
            assert_same(model)
    return lens, took


def test_step_async_rides_behind_a_slow_pass(mock_backend):
    """A slow driver (5 ms per map / set_access) makes the eager pass take tens of milliseconds.  The
    next decode steps need no new page, so they are queued behind it and return at once; the
    bookkeeping ends up exactly where the oracle's strictly sequential replay ends up."""
    model, _ = make_pair(2, 2, 64, 4, 32768, mem_pages=64)
    tpp = model.tokens_per_page
    lens = [tpp - 9, tpp - 30, 100, 5000]            # the first pass looks 9 tokens ahead: no new page yet
    va.step_async(lens)
    model.step_async(lens)
    va.wait_background()
    va.mock_set_call_delay_us(5000)
    lens, took = _decode_steps(model, lens, 3, 32768)
    st = va.get_step_stats()                         # waits for the mapper
    # pass of step 1 maps request 0's next page in both layers (4 maps + 4 set_access = 40 ms); the
    # two steps after it rode behind
    assert st["queued_steps"] >= 1
    assert st["max_background_ns"] > 30e6
    assert min(took[1:]) < 0.020, took         # (the pass it rides behind takes >= 40 ms)
    assert st["total_sync_pages"] == 2 * 2 * 4       # only the very first step mapped on the critical path
    assert_same(model)
    va.mock_set_call_delay_us(0)
    lens, _ = _decode_steps(model, lens, 40, 32768, read_state_every=7)   # across the boundary
    assert_same(model)
    assert va.get_state()["mapped_pages"][0] == 2


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("nodefer", [False, True])
def test_queued_steps_match_oracle_on_random_traces(mock_backend, seed, nodefer):
    """Random serving traces with a slow driver and NO state reads between decode steps: whichever
    steps happen to be queued, mapped_pages / pool order / page map / driver calls equal the oracle's."""
    L, Hkv, D, B, ctx = 2, 2, 64, 5, 32768
    model, _ = make_pair(L, Hkv, D, B, ctx, mem_pages=200)
    if nodefer:
        va.set_deferred_reclamation(False)
        model.set_deferred_reclamation(False)
    tpp = model.tokens_per_page
    rng = random.Random(500 + seed)
    va.mock_set_call_delay_us(150)
    lens = [0] * B
    for it in range(60):
        r = rng.random()
        active = [i for i in range(B) if lens[i]]
        if r < 0.3 and len(active) < B:
            n = rng.choice([1, tpp - 2, tpp, tpp + 1, 2 * tpp - 3, rng.randrange(1, ctx // 3)])
            got, want = va.alloc_new_batch_idx(n), model.alloc_new_batch_idx(n)
            assert got == want
            if got >= 0:
                lens[got] = n
        elif r < 0.4 and active:
            i = rng.choice(active)
            va.free_batch_idx(i)
            model.free_batch_idx(i)
            lens[i] = 0
        try:
            lens, _ = _decode_steps(model, lens, rng.randrange(1, 6), ctx)
        except (RuntimeError, AllocatorOOM):
            pytest.skip("trace ran out of pages")   # (sized not to; OOM parity is covered above)
    assert_same(model)
    st = va.get_step_stats()
    assert st["steps"] > 60
    # ... and the driver saw the maps / unmaps in the reference's order with the reference's pages
    log = va.get_driver_log()
    reserves = [r for r in log if r[0] == 1]
    base = {f"k{i}": reserves[i][1] for i in range(L)}
    base.update({f"v{i}": reserves[L + i][1] for i in range(L)})
    got = [(r[0], r[1], r[3] - 1 if r[0] == 3 else -1) for r in log if r[0] in (3, 5)]
    want = [(3 if c.op == "map" else 5, base[c.tensor] + c.offset, c.page if c.op == "map" else -1)
            for c in model.calls if c.op in ("map", "unmap")]
    assert got == want


def test_queueing_is_refused_when_the_pass_takes_pages_back(mock_backend):
    """With the pool empty the background pass takes an inactive request's page back
    (reclaim_on_demand) to map the page request 0 needs next: no step may be queued behind THAT pass,
    the call waits like the reference does.  Either way the books equal the oracle's."""
    model, _ = make_pair(1, 2, 64, 3, 32768, mem_pages=2 * 3)   # three blocks in all
    tpp = model.tokens_per_page
    lens = [tpp - 5, 10, 10]
    va.step_async(lens)
    model.step_async(lens)
    va.free_batch_idx(2)                                         # keeps its page (deferred reclamation)
    model.free_batch_idx(2)
    lens[2] = 0
    assert_same(model)
    va.mock_set_call_delay_us(2000)
    lens, _ = _decode_steps(model, lens, 9, 32768)               # request 0 crosses the boundary
    assert_same(model)
    st = va.get_state()
    assert st["mapped_pages"] == [2, 1, 0]                       # the page moved from request 2 to request 0
    assert va.get_step_stats()["queued_steps"] < 9
    va.mock_set_call_delay_us(0)


def test_set_queueing_off_waits_like_the_reference(mock_backend):
    model, _ = make_pair(2, 2, 64, 4, 32768, mem_pages=64)
    tpp = model.tokens_per_page
    va.set_queueing(False)
    lens = [tpp - 9, tpp - 30, 100, 5000]
    va.step_async(lens)
    model.step_async(lens)
    va.wait_background()
    va.mock_set_call_delay_us(3000)
    lens, took = _decode_steps(model, lens, 3, 32768)
    assert va.get_step_stats()["queued_steps"] == 0
    assert took[1] > 0.015                           # stood still for the pass of the step before
    assert_same(model)
    va.mock_set_call_delay_us(0)


def test_queued_step_records_its_fence_in_the_other_slot(mock_backend):
    """Unmaps wait for an event recorded on the compute stream at THEIR step's step_async.  A queued
    step records into the slot the pass in flight is not using; the mapper switches slots when it
    takes the step on."""
    model, _ = make_pair(1, 2, 64, 3, 32768, mem_pages=12)
    tpp = model.tokens_per_page
    va.set_compute_stream(0, True)
    lens = [tpp - 9, 7, 0]
    va.step_async(lens)
    model.step_async(lens)
    va.wait_background()
    va.mock_set_call_delay_us(4000)
    lens, _ = _decode_steps(model, lens, 3, 32768)
    st = va.get_step_stats()
    fc = va.mock_fence_counts()
    assert sum(fc["records"]) == st["steps"]
    if st["queued_steps"]:
        assert fc["records"][0] and fc["records"][1]
    assert_same(model)
    va.mock_set_call_delay_us(0)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("nodefer", [False, True])
def test_queued_steps_with_slots_switched_by_the_length_vector(mock_backend, seed, nodefer):
    """Slots are activated and dropped through the length vector alone (no alloc / free call in
    between, so nothing but step_async ever meets the mapper): a step that re-activates a slot the pass
    in flight may be taking pages from, or that needs a page nobody mapped yet, must wait; every
    other step may ride behind.  The books must equal the oracle's at every checkpoint."""
    L, Hkv, D, B, ctx = 2, 2, 64, 6, 32768
    model, _ = make_pair(L, Hkv, D, B, ctx, mem_pages=4 * 14)          # 14 blocks: reclaim happens
    if nodefer:
        va.set_deferred_reclamation(False)
        model.set_deferred_reclamation(False)
    tpp = model.tokens_per_page
    rng = random.Random(900 + seed)
    va.mock_set_call_delay_us(120)
    lens = [0] * B
    for it in range(220):
        for i in range(B):
            r = rng.random()
            if lens[i] == 0:
                if r < 0.06:
                    lens[i] = rng.choice([1, tpp - 3, tpp, tpp + 1, 2 * tpp - 2, rng.randrange(1, 3 * tpp)])
            elif r < 0.04:
                lens[i] = 0
            else:
                lens[i] = min(lens[i] + rng.choice([1, 1, 1, 1, 7]), ctx - 1)
        err_got = err_want = None
        try:
            va.step_async(lens)
        except RuntimeError as e:
            err_got = str(e)
        try:
            model.step_async(lens)
        except AllocatorOOM as e:
            err_want = str(e)
        assert (err_got is None) == (err_want is None), (it, err_got, err_want)
        if err_want:
            va.set_verbose(False)
            big = max(range(B), key=lambda i: lens[i])
            lens[big] = 0
            assert_same(model)
        elif rng.random() < 0.1:
            assert_same(model)
    assert_same(model)
