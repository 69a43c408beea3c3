"""Host logic on CPU: the cache-engine mirror against the allocator (mock driver) + oracle, the
three wrapper mirrors with oracle-backed operators injected, and head-sharded TP under gloo
(world_size 2)."""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attention_ref as ref
from oracle.allocator_model import MB, AllocatorModel
from vattention_b200 import _lib
from vattention_b200 import vattention as va
from vattention_b200.cache_engine import vATTNCacheEngine
from vattention_b200.tp import HeadShard, HeadShardedAttention
from vattention_b200.wrappers import (VAttentionFlashAttentionPODWrapper, VAttentionFlashAttentionWrapper,
                                      VAttentionFlashInferWrapper, get_attention_wrapper_class)


class Seq:
    def __init__(self, seq_id, prompt_len):
        self.seq_id, self.prompt_len, self.processed, self.generated = seq_id, prompt_len, 0, 0
        self.finished = False

    def get_next_prompt_chunk_len(self, chunk):
        return min(chunk, self.prompt_len - self.processed)

    def get_num_prompt_tokens_processed(self):
        return self.processed

    def get_len(self):
        return self.prompt_len + self.generated

    def is_finished(self):
        return self.finished


class MD:
    def __init__(self, seq, is_prompt, chunk=0):
        self.seq, self.is_prompt, self.prompt_chunk_len = seq, is_prompt, chunk


# oracle-backed operator namespace with the product's call surface
def _ops():
    def flash_attn_with_kvcache(q, kc, vc, k=None, v=None, cache_seqlens=None, cache_batch_idx=None,
                                block_table=None, softmax_scale=None, causal=False, **kw):
        return ref.attn_with_kvcache_ref(q, kc, vc, k, v, cache_seqlens, cache_batch_idx, softmax_scale, causal)

    def single_prefill_with_kv_cache(q, k, v, causal=False, **kw):
        return ref.single_prefill_ref(q, k, v, causal)

    def true_fused_attn_with_kvcache(q_p, kc_p, vc_p, q_d, kc_d, vc_d, k=None, v=None, causal=False,
                                     cache_seqlens_p=None, cache_seqlens_d=None, cache_batch_idx=None,
                                     softmax_scale=None, fused_params=15, **kw):
        o_p = o_d = None
        if q_p is not None:
            o_p = ref.attn_with_kvcache_ref(q_p, kc_p, vc_p, cache_seqlens=cache_seqlens_p,
                                            softmax_scale=softmax_scale, causal=causal)
        if q_d is not None:
            o_d = ref.attn_with_kvcache_ref(q_d, kc_d, vc_d, k, v, cache_seqlens_d, cache_batch_idx,
                                            softmax_scale, causal)
        return o_p, o_d

    def cache_flat(key, value, kc, vc, dt):
        ref.cache_flat_ref(key, value, kc, vc)

    return types.SimpleNamespace(flash_attn_with_kvcache=flash_attn_with_kvcache,
                                 single_prefill_with_kv_cache=single_prefill_with_kv_cache,
                                 true_fused_attn_with_kvcache=true_fused_attn_with_kvcache,
                                 cache_flat=cache_flat)


@pytest.fixture()
def mock_backend():
    va._use_backend(_lib.BACKEND_HOST_MOCK)
    yield
    va.cleanup()
    va._use_backend(_lib.BACKEND_CUDA)


def test_cache_engine_call_pattern_matches_oracle(mock_backend):
    L, Hkv, D, B, ctx = 2, 2, 64, 4, 16384
    wrapper = VAttentionFlashAttentionWrapper(ops=_ops()).init(num_q_heads=4, num_kv_heads=Hkv, head_dim=D,
                                                               device=torch.device("cpu"))
    eng = vATTNCacheEngine(L, Hkv, D, B, ctx, torch.float16, 2 * MB, 64 * MB, torch.device("cpu"),
                           "async", False, wrapper)
    model = AllocatorModel(L, Hkv, D, B, ctx, 2, 2 * MB, False)
    model.reserve_physical_pages(64 * MB)
    assert len(eng.gpu_cache) == L and tuple(eng.gpu_cache[0][0].shape) == (B, ctx, Hkv, D)
    s0, s1, s2 = Seq(10, 9000), Seq(11, 300), Seq(12, 5)
    # iteration 1: s0 prefill chunk 4096, s1 full prompt
    eng.step([MD(s0, True, 4096), MD(s1, True, 4096)])
    assert model.alloc_new_batch_idx(4096) == eng.seq_to_batch_idx[10] == 0
    assert model.alloc_new_batch_idx(300) == eng.seq_to_batch_idx[11] == 1
    model.step_async([4096, 300, 0, 0])
    assert eng.curr_batch_idx.tolist() == [0, 1] and wrapper.batch_index_gen.numel() == 0
    s0.processed, s1.processed, s1.generated = 4096, 300, 1
    # iteration 2: s1 decodes, s0 next chunk, s2 arrives -> batch idx = prefills then decodes
    eng.step([MD(s1, False), MD(s0, True, 4096), MD(s2, True, 4096)])
    assert model.alloc_new_batch_idx(5) == eng.seq_to_batch_idx[12] == 2
    model.step_async([8192, 301, 5, 0])
    assert eng.curr_batch_idx.tolist() == [0, 2, 1] and wrapper.batch_index_gen.tolist() == [1]
    va.wait_background()
    st = va.get_state()
    assert st["mapped_pages"] == model.mapped_pages and st["seq_lens"] == model.seq_lens
    assert eng.num_free_blocks() == model.num_free_kvblocks()
    # s1 finishes: its slot is freed but pages stay mapped (deferred reclamation)
    s1.finished = True
    eng.on_step_completion([MD(s1, False)])
    model.free_batch_idx(1)
    assert 11 not in eng.seq_to_batch_idx and eng.curr_seq_lens[1] == 0
    assert va.get_state()["mapped_pages"] == model.mapped_pages
    # a new short sequence reuses the freed slot with its mapped page (best fit)
    s3 = Seq(13, 100)
    eng.step([MD(s3, True, 4096)])
    assert eng.seq_to_batch_idx[13] == model.alloc_new_batch_idx(100) == 1
    with pytest.raises(Exception, match="not found in req_table"):
        eng.free_request(999)
    eng.preempt_requests([s0])
    assert 10 not in eng.seq_to_batch_idx
    eng.reclaim_req_ids()
    assert eng.seq_to_batch_idx == {} and eng.curr_seq_lens == [0] * B


def _mixed_batch(dtype=torch.float32):
    """One prefill chunk (64 cached + 32 new) and three decodes, as flat [tokens, H*D] tensors."""
    g = torch.Generator().manual_seed(0)
    Hq, Hkv, D, B, ctx = 4, 2, 64, 5, 256
    kc = torch.randn(B, ctx, Hkv, D, generator=g).to(dtype)
    vc = torch.randn(B, ctx, Hkv, D, generator=g).to(dtype)
    p = Seq(1, 200)
    p.processed = 64
    d1, d2, d3 = Seq(2, 50), Seq(3, 120), Seq(4, 7)
    for s, gen in ((d1, 3), (d2, 1), (d3, 9)):
        s.processed, s.generated = s.prompt_len, gen
    mds = [MD(d1, False), MD(p, True, 32), MD(d2, False), MD(d3, False)]
    ntok = 32 + 3
    q = torch.randn(ntok, Hq * D, generator=g).to(dtype)
    k = torch.randn(ntok, Hkv * D, generator=g).to(dtype)
    v = torch.randn(ntok, Hkv * D, generator=g).to(dtype)
    slots = {1: 3, 2: 0, 3: 4, 4: 1}
    return (Hq, Hkv, D), (kc, vc), mds, (q, k, v), slots, (p, [d1, d2, d3])


@pytest.mark.parametrize("cls", [VAttentionFlashAttentionWrapper, VAttentionFlashInferWrapper,
                                 VAttentionFlashAttentionPODWrapper])
def test_wrappers_assemble_the_batch_like_the_reference(cls):
    (Hq, Hkv, D), (kc, vc), mds, (q, k, v), slots, (p, decs) = _mixed_batch()
    scale = D ** -0.5
    w = cls(ops=_ops()).init(num_q_heads=Hq, num_kv_heads=Hkv, head_dim=D, device=torch.device("cpu"))
    b_idx = torch.tensor([slots[p.seq_id]] + [slots[s.seq_id] for s in decs], dtype=torch.int32)
    w.set_batch_idx(b_idx, b_idx[1:])
    w.begin_forward(mds)
    assert w.prefill_query_lens == [32] and w.prefill_cache_lens == [64]
    assert w.decode_cache_lens.tolist() == [s.get_len() - 1 for s in decs]
    assert w.max_cache_len == max(s.get_len() for s in decs)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    out = w.forward(q, k, v, (kc, vc), softmax_scale=scale, layer_id=0)
    w.end_forward()
    assert w.batch_index is None and not w.is_metadata_initialized
    # expected, sequence by sequence
    sp = slots[p.seq_id]
    kc_ref[sp, 64:96] = k[:32].view(32, Hkv, D)
    vc_ref[sp, 64:96] = v[:32].view(32, Hkv, D)
    want_p = ref.attn_with_kvcache_ref(q[:32].view(1, 32, Hq, D), kc_ref[sp:sp + 1], vc_ref[sp:sp + 1],
                                       cache_seqlens=torch.tensor([96], dtype=torch.int32),
                                       softmax_scale=scale, causal=True)
    assert torch.allclose(out[:32], want_p.reshape(32, -1), atol=1e-5)
    for j, s in enumerate(decs):
        sl = slots[s.seq_id]
        L0 = s.get_len() - 1
        kc_ref[sl, L0] = k[32 + j].view(Hkv, D)
        vc_ref[sl, L0] = v[32 + j].view(Hkv, D)
        want = ref.attn_with_kvcache_ref(q[32 + j].view(1, 1, Hq, D), kc_ref[sl:sl + 1], vc_ref[sl:sl + 1],
                                         cache_seqlens=torch.tensor([L0 + 1], dtype=torch.int32),
                                         softmax_scale=scale)
        assert torch.allclose(out[32 + j], want.reshape(-1), atol=1e-5)
    assert torch.equal(kc, kc_ref) and torch.equal(vc, vc_ref)  # chunk + appended tokens landed in the cache


def test_wrapper_registry_and_profiling_short_circuit():
    assert get_attention_wrapper_class("FA_VATTN") is VAttentionFlashAttentionWrapper
    assert get_attention_wrapper_class("fi_vattn_sync") is VAttentionFlashInferWrapper
    assert get_attention_wrapper_class("fa_pod_megacache") is VAttentionFlashAttentionPODWrapper
    with pytest.raises(ValueError):
        get_attention_wrapper_class("flash_attention")       # paged backends are out of scope
    w = VAttentionFlashAttentionWrapper(ops=_ops()).init(num_q_heads=2, num_kv_heads=2, head_dim=64,
                                                         device=torch.device("cpu"))
    w.begin_forward([])
    w.is_profiling_iteration = True                           # model_runner.py:194-195
    x = torch.randn(3, 128)
    assert torch.all(w.forward(x, x, x, None) == 0)
    pod = VAttentionFlashAttentionPODWrapper(ops=_ops()).init(num_q_heads=2, num_kv_heads=2, head_dim=64,
                                                              device=torch.device("cpu"))
    a, b = Seq(1, 10), Seq(2, 10)
    with pytest.raises(ValueError, match="Batched prefills"):
        pod.begin_forward([MD(a, True, 10), MD(b, True, 10)])


def test_head_shard_arithmetic():
    sh = HeadShard(rank=3, world=8, num_heads=64, num_kv_heads=8, head_dim=128)   # Llama-3-70B TP8
    assert sh.heads_per_rank == 8 and sh.kv_heads_per_rank == 1
    assert list(sh.q_range) == list(range(24, 32)) and list(sh.kv_range) == [3]
    assert sh.tokens_per_page(2 * MB) == 8192
    with pytest.raises(ValueError):
        HeadShard(0, 3, 32, 8, 128)
    w = torch.arange(64 * 128 * 4, dtype=torch.float32).view(64 * 128, 4)
    assert torch.equal(sh.shard_o_proj(w), w[24 * 128:32 * 128])


def _tp_worker(rank, world, port, q_out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                       # identical "weights" and inputs on every rank
    Hq, Hkv, D, B, S, hidden = 8, 4, 64, 3, 96, 32
    q = torch.randn(B, 1, Hq, D)
    kc, vc = torch.randn(B, S, Hkv, D), torch.randn(B, S, Hkv, D)
    kn, vn = torch.randn(B, 1, Hkv, D), torch.randn(B, 1, Hkv, D)
    lens = torch.tensor([95, 10, 40], dtype=torch.int32)
    w_o = torch.randn(Hq * D, hidden) * 0.1
    sh = HeadShard(rank, world, Hq, Hkv, D)
    attn = lambda qs, kcs, vcs, ks, vs, **kw: ref.attn_with_kvcache_ref(qs, kcs, vcs, ks, vs, **kw)
    tp = HeadShardedAttention(sh, sh.shard_o_proj(w_o), attn)
    got = tp.forward(sh.shard_q(q), sh.shard_kv(kc).clone(), sh.shard_kv(vc).clone(), sh.shard_kv(kn),
                     sh.shard_kv(vn), cache_seqlens=lens, causal=True)
    full = ref.attn_with_kvcache_ref(q, kc.clone(), vc.clone(), kn, vn, lens, causal=True)
    want = full.reshape(B, -1) @ w_o
    q_out.put((rank, float((got - want).abs().max())))
    dist.destroy_process_group()


def test_head_sharded_attention_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_tp_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q_out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(err < 1e-4 for _, err in res), res


def test_block_space_manager_accounting_against_allocator():
    """vattention_block_space_manager.py:9-98 mirror, fed from the allocator (mock driver): what the
    scheduler admits never exceeds what step() can map."""
    import torch
    from vattention_b200 import _lib
    from vattention_b200 import vattention as va
    from vattention_b200.block_space_manager import vAttentionBlockSpaceManager

    class Seq:
        def __init__(self, i, n):
            self.seq_id, self.n = i, n

        def get_len(self):
            return self.n

    va._use_backend(_lib.BACKEND_HOST_MOCK)
    try:
        L, B, ctx = 2, 4, 32768
        va.init_kvcache(L, 2, 64, B, ctx, 0, torch.float16, 2 << 20, False)
        va.reserve_physical_pages(20 * (2 << 20))          # 20 pages = 5 blocks of 2L pages
        tpp = va.get_config()["tokens_per_page"]
        m = vAttentionBlockSpaceManager(tpp, 5, ctx, watermark=0.2)
        assert m.watermark_blocks == 1
        with pytest.raises(AttributeError):
            m.can_append_slot()                              # free_blocks unset, as in the reference
        assert m.refresh(va) == 5
        a, b, c = Seq(0, 2 * tpp), Seq(1, tpp + 1), Seq(2, 1)
        assert m.get_num_blocks(a) == 2 and m.get_num_blocks(b) == 2 and m.get_num_blocks(c) == 1
        assert m.can_allocate(a)
        m.allocate(a)
        assert m.promised_blocks == 2 and m.is_allocated(a) and not m.is_allocated(b)
        assert m.can_allocate(b)                             # 5 - 2 - 2 >= 1
        m.allocate(b)
        assert not m.can_allocate(c)                         # 5 - 4 - 1 < 1
        assert m.can_append_slot()
        # the allocator maps exactly what was promised
        lens = [0] * B
        for s in (a, b):
            lens[va.alloc_new_batch_idx(s.n)] = s.n
        va.step(lens, True)
        assert m.refresh(va) == 1 and m.promised_blocks == 0
        assert not m.can_append_slot() or m.free_blocks == 1
        m.append_slot(a)                                     # 2*tpp -> 2*tpp+1 opens block 3
        assert m.promised_blocks == 1 and not m.can_append_slot()
        m.append_slot(b)                                     # tpp+1 -> tpp+2 stays inside block 2
        assert m.promised_blocks == 1
        m.free(a)
        m.free(a)                                            # second free is a no-op
        assert m.free_blocks == 3 and not m.is_allocated(a)
        va.free_batch_idx(0)
        assert m.refresh(va) == va.num_free_kvblocks() == 3  # deferred reclaim: a's pages count as free
        assert m.get_block_table(b) is None and m.get_num_free_gpu_blocks(b) == 3
        m.reset()
        assert not m.is_allocated(b)
    finally:
        va.cleanup()
        va._use_backend(_lib.BACKEND_CUDA)


def test_bench_crossing_schedule_is_a_fixed_fraction_of_the_page_at_every_rank_count():
    """bench.py's start lengths: sequences cross the page boundary one after another, the window
    they are spread over is the same fraction of a rank's page for N = 1, 2, 4, 8, and at least one
    sequence maps a page inside the timed steps."""
    import bench
    page, D = 2 << 20, 128
    for batch, hkv_full in ((64, 8), (16, 8)):
        fractions = set()
        for world in (1, 2, 4, 8):
            hkv = hkv_full // world
            tpp, tpp_full = page // (hkv * D * 2), page // (hkv_full * D * 2)
            W, K, ctx = 3, 8, 32768
            total = 3 + W + K + K + W + K + 2
            spread, lens = bench.crossing_schedule(batch, ctx, tpp, tpp_full, W, K, total)
            assert len(lens) == batch and ctx % tpp == 0
            fractions.add(tpp / spread)
            assert all(ctx - spread - (W + 2) - 1 <= n < ctx for n in lens)
            # steps 1.. of the run advance every length by one; the timed steps are W + 3 .. W + 2 + K
            crossing_steps = [ctx - n + 1 for n in lens]          # the step whose token is the page's first
            timed = [s for s in crossing_steps if W + 3 <= s <= W + 2 + K]
            assert timed, (batch, world, sorted(crossing_steps)[:4])
            assert len(timed) <= max(1, -(-batch * K // spread) + 1)
        assert len(fractions) == 1, fractions
