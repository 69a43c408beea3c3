"""CPU restatement of the reference vAttention allocator's bookkeeping.

TEST INFRASTRUCTURE ONLY.  Nothing under vattention_b200/ imports this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.

It restates, function by function, the integer arithmetic and container
discipline of /root/reference/vattention (commit ef3fff25):

    vattention.cu   policy (init sizes, grow, step, reclaim, reqId best-fit)
    utils.h         helpers (tokens_to_pages, need_new_page_async, overcommit)
    mux.h           page pool LIFO + MAP/UNMAP macros
    cudaInternal.h  reserve_cuda_pages, map_cuda_pages call order

so that a trace of API calls yields, bit for bit, the same `mapped_pages`,
`curr_seq_lengths`, free-pool order, page map and return values as the
reference.  Physical page handles are replaced by their 0-based creation index
(the reference's pool after reserve_cuda_pages is in creation order,
cudaInternal.h:51-56), virtual addresses by (tensor, offset).

The background thread (vattention.cu:538-546) is run to completion inside
step_async: that is the schedule in which the thread finishes before the next
API call, which is also the only schedule the new implementation allows.

Parity pinning: the reference ships no tests or golden vectors for the
allocator.  tests/golden/alloc_trace_*.json are outputs of the
reference itself (vattention.cu compiled unmodified by oracle/Makefile and run on
a B200 by oracle/gen_alloc_golden.py); tests/test_oracle_golden.py checks this
model against them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

U64 = (1 << 64) - 1
MB = 1024 * 1024

EAGER_NUM_STEPS = 10   # vattention.cu:486
EAGER_NUM_KVBLOCKS = 2  # vattention.cu:487

OOM_MSG = "***** OOM on demand: not enough free pages to continue *****"
POOL_EMPTY_MSG = "***** page pool is empty *****"


class AllocatorOOM(RuntimeError):
    pass


@dataclass
class DriverCall:
    op: str       # create | map | set_access | unmap | reserve
    tensor: str   # "k3" / "v0" / ""
    offset: int
    page: int     # creation index, -1 when n/a


@dataclass
class AllocatorModel:
    num_layers: int
    num_kv_heads: int
    head_size: int
    max_batch_size: int
    max_context_length: int
    bytes_per_elem: int
    page_size: int = 2 * MB
    megacache: bool = False

    deferred_reclaim: bool = True   # utils.h:78
    pool: List[int] = field(default_factory=list)
    pagemap: Dict[Tuple[int, int, int], Tuple[int, int]] = field(default_factory=dict)
    mapped_pages: List[int] = field(default_factory=list)
    seq_lens: List[int] = field(default_factory=list)
    calls: List[DriverCall] = field(default_factory=list)
    created: int = 0

    def __post_init__(self) -> None:
        # vattention.cu:38-47 init_kv_block_size
        row = self.num_kv_heads * self.head_size * self.bytes_per_elem
        if self.megacache:
            row *= self.num_layers
        self.tokens_per_page = self.page_size // row
        # vattention.cu:49-74 init_buffer_sizes
        self.virt_buff_size_per_token = row
        per_req = row * self.max_context_length
        per_req = (per_req + self.page_size - 1) // self.page_size * self.page_size
        self.max_pages_per_req = per_req // self.page_size
        rem = per_req % self.page_size
        if rem:
            per_req += self.page_size - rem
        self.virt_buff_size_per_req = per_req
        self.virt_buff_size = per_req * self.max_batch_size
        # utils.h:88-97 init_kvcache_batch_metadata
        self.mapped_pages = [0] * self.max_batch_size
        self.seq_lens = [0] * self.max_batch_size
        self.num_tensors_per_side = 1 if self.megacache else self.num_layers

    # ---------------------------------------------------------------- utils.h
    def tokens_to_pages(self, n: int) -> int:  # utils.h:105-108
        return (n + self.tokens_per_page - 1) // self.tokens_per_page

    def _blocks_in_pool(self) -> int:  # utils.h:8-11
        if self.megacache:
            return len(self.pool) // 2
        return len(self.pool) // (2 * self.num_layers)

    def kvblocks_available(self, n: int) -> bool:  # vattention.cu:212-217
        return self._blocks_in_pool() >= n

    def _overcommitted(self) -> int:  # utils.h:177-183 (u64 wrap-around kept)
        acc = 0
        for r in range(self.max_batch_size):
            acc = (acc + self.mapped_pages[r] - self.tokens_to_pages(self.seq_lens[r])) & U64
        return acc

    def need_new_page_async(self, req: int, eager: int) -> int:  # utils.h:206-219
        if self.seq_lens[req] == 0:
            return 0
        have = self.mapped_pages[req]
        if have == self.max_pages_per_req:
            return 0
        need = self.tokens_to_pages(self.seq_lens[req] + eager)
        return 0 if need <= have else need - have

    # ----------------------------------------------------------------- mux.h
    def _pop(self) -> int:  # mux.h:1-8
        if not self.pool:
            raise AllocatorOOM(POOL_EMPTY_MSG)
        return self.pool.pop()

    def _map_pages(self, req: int, layer: int, off: int) -> None:
        # mux.h:37-48 MAP_PAGES + cudaInternal.h:70-82 map_cuda_pages
        k = self._pop()
        v = self._pop()
        self.calls.append(DriverCall("map", f"k{layer}", off, k))
        self.calls.append(DriverCall("map", f"v{layer}", off, v))
        self.calls.append(DriverCall("set_access", f"k{layer}", off, -1))
        self.calls.append(DriverCall("set_access", f"v{layer}", off, -1))
        self.pagemap[(req, off, layer)] = (k, v)

    def _unmap_pages(self, req: int, layer: int, off: int) -> None:
        # mux.h:51-66 UNMAP_PAGES
        self.calls.append(DriverCall("unmap", f"k{layer}", off, -1))
        self.calls.append(DriverCall("unmap", f"v{layer}", off, -1))
        k, v = self.pagemap.pop((req, off, layer))
        self.pool.append(k)
        self.pool.append(v)

    # ------------------------------------------------------- cudaInternal.h
    def reserve_physical_pages(self, free_memory: int) -> int:
        # utils.h:221-228 get_num_phys_blocks, cudaInternal.h:45-59
        n = free_memory // self.page_size
        n -= n % (2 * self.num_layers)
        while len(self.pool) < n:
            self.calls.append(DriverCall("create", "", 0, self.created))
            self.pool.append(self.created)
            self.created += 1
        return len(self.pool)

    # --------------------------------------------------------- vattention.cu
    def num_free_kvblocks(self) -> int:  # :194-211
        return (self._blocks_in_pool() + self._overcommitted()) & U64

    def _unmap_req_page_one(self, req: int) -> None:  # :219-241, utils.h:193-204
        assert self.mapped_pages[req] > 0
        off = req * self.virt_buff_size_per_req + (self.mapped_pages[req] - 1) * self.page_size
        for layer in range(self.num_tensors_per_side):
            self._unmap_pages(req, layer, off)
        self.mapped_pages[req] -= 1

    def _release_some(self, req: int, retain: int) -> None:  # :243-252
        while self.mapped_pages[req] > retain:
            self._unmap_req_page_one(req)

    def _grow(self, req: int, num_blocks: int, sync: bool) -> None:  # :268-323
        if num_blocks <= 0:
            return
        if not self.kvblocks_available(num_blocks):
            if not sync:
                return
            raise AllocatorOOM(OOM_MSG)
        for _ in range(num_blocks):
            off = req * self.virt_buff_size_per_req + self.mapped_pages[req] * self.page_size
            if not off < (req + 1) * self.virt_buff_size_per_req:  # :254-266
                return
            for layer in range(self.num_tensors_per_side):
                self._map_pages(req, layer, off)
            self.mapped_pages[req] += 1

    def _reclaim_on_demand(self, num_kvblocks: int) -> None:  # :420-438
        for req in range(self.max_batch_size - 1, -1, -1):
            if self.kvblocks_available(num_kvblocks):
                break
            have = self.mapped_pages[req]
            need = self.tokens_to_pages(self.seq_lens[req])
            if have <= need:
                continue
            self._release_some(req, need)

    def _map_pages_for_curr_step(self, req: int, seq_len: int) -> None:  # :376-392
        need = self.tokens_to_pages(seq_len)
        have = self.mapped_pages[req]
        if need <= have:
            return
        need -= have
        if not self.kvblocks_available(need):
            self._reclaim_on_demand(need)
        self._grow(req, need, True)
        self.seq_lens[req] = seq_len

    def step(self, seq_lens: List[int], eager_reclaim: bool) -> None:  # :395-409 step_sync
        for req in range(self.max_batch_size):
            self.seq_lens[req] = seq_lens[req]
            if eager_reclaim and seq_lens[req] == 0 and self.mapped_pages[req] != 0:
                self._release_some(req, 0)
                continue
            self._map_pages_for_curr_step(req, seq_lens[req])

    def _do_reclaim_pages(self) -> None:  # :444-469
        if self.deferred_reclaim:
            return
        next_prefill = -1
        for req in range(self.max_batch_size):
            if self.seq_lens[req] == 0:
                next_prefill = req
                break
        for req in range(self.max_batch_size - 1, -1, -1):
            if self.seq_lens[req] != 0 or req == next_prefill:
                continue
            if self.mapped_pages[req] == 0:
                continue
            self._unmap_req_page_one(req)
            break

    def _do_kvcache_memory_management(self) -> None:  # :488-536
        nr_required = 0
        nr_mapped_curr = 0
        done = False
        for req in range(self.max_batch_size):
            nr_required += self.need_new_page_async(req, 1)
        if not self.kvblocks_available(nr_required):
            self._reclaim_on_demand(nr_required)
        if not self.kvblocks_available(nr_required):
            return
        eager = 1
        while eager < EAGER_NUM_STEPS and not done:
            for req in range(self.max_batch_size):
                n = self.need_new_page_async(req, eager)
                self._grow(req, n, False)
                nr_mapped_curr += n
                if eager == 1:
                    continue
                if nr_mapped_curr >= EAGER_NUM_KVBLOCKS:
                    done = True
                    break
            eager += 1
        if nr_required:
            return
        self._do_reclaim_pages()

    def step_async(self, seq_lens: List[int]) -> None:  # :549-558
        self.seq_lens = list(seq_lens)
        for req in range(self.max_batch_size):  # prepare_prefill_kvcache :412-418
            self._map_pages_for_curr_step(req, self.seq_lens[req])
        self._do_kvcache_memory_management()    # spawn_kvcache_manager, run to completion

    def alloc_new_batch_idx(self, seqlen: int) -> int:  # :564-589
        new_id = -1
        need = self.tokens_to_pages(seqlen)
        for req in range(self.max_batch_size):
            if self.seq_lens[req] != 0:
                continue
            if new_id == -1:
                new_id = req
                continue
            if self.mapped_pages[req] >= need and self.mapped_pages[req] < self.mapped_pages[new_id]:
                new_id = req
        if new_id != -1:
            self.seq_lens[new_id] = seqlen
        return new_id

    def free_batch_idx(self, req: int) -> None:  # :591-594
        self.seq_lens[req] = 0

    def set_deferred_reclamation(self, val: bool) -> None:  # :596-599
        self.deferred_reclaim = val

    def release_all(self) -> None:  # DO_KVCACHE_CLEANUP, mux.h:24-35 (page part)
        for req in range(self.max_batch_size):
            self._release_some(req, 0)

    # ------------------------------------------------------------ snapshots
    def snapshot(self) -> dict:
        return {
            "mapped_pages": list(self.mapped_pages),
            "seq_lens": list(self.seq_lens),
            "pool": list(self.pool),
            "pagemap": sorted([list(k) + list(v) for k, v in self.pagemap.items()]),
            "num_free_kvblocks": self.num_free_kvblocks(),
        }


def page_size_to_block_tokens(page_size: int, num_kv_heads: int, head_dim: int, tp: int,
                              num_layers: int, megacache: bool, itemsize: int = 2) -> int:
    """sarathi-lean/sarathi/engine/arg_utils.py:147-159: how the engine converts
    `model_block_size` (page bytes) into tokens per scheduler block."""
    # the reference chains floor divisions in this order; keep it (it differs from
    # page // (Hkv/tp * D * L * itemsize) when the intermediate quotients truncate)
    block = page_size // (num_kv_heads // tp)
    block = block // head_dim
    if megacache:
        block = block // num_layers
    return block // itemsize
