"""Host-side mirror of sarathi's vATTNCacheEngine -- the allocator's only in-tree caller and
therefore its contract (sarathi-lean/sarathi/worker/cache_engine/vATTN_cache_engine.py:18-195).

Keeps seq_id -> reqId (batch slot) and the per-slot `curr_seq_lens` vector, asks the allocator
for a slot when a sequence first appears, passes the whole vector to step_async / step once per
iteration, builds the batch-index tensors the attention wrapper needs (prefills first, then
decodes), and frees slots of finished / preempted sequences.  Sequence metadata is duck-typed
(see wrappers.py).  The allocator module is injectable so the logic runs on the mock driver.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import vattention as _default_allocator


class vATTNCacheEngine:
    def __init__(self, num_layers: int, num_kv_heads: int, head_size: int, max_batch_size: int,
                 max_model_seq_len: int, dtype: torch.dtype, page_size: int, memory_for_gpu: int,
                 device: torch.device, mem_alloc_backend: str = "async", megacache: bool = False,
                 attention_wrapper=None, allocator=None):
        self.alloc = allocator or _default_allocator
        self.num_layers, self.num_heads, self.head_size = num_layers, num_kv_heads, head_size
        self.max_batch_size = max_batch_size
        self.max_model_seq_len = max_model_seq_len
        self.dtype, self.page_size = dtype, page_size
        self.device = device
        self.device_idx = device.index or 0 if device.type == "cuda" else 0
        self.vattn_async = mem_alloc_backend == "async"      # cache_engine/__init__.py:20-25
        self.vattn_mega_cache = megacache
        self.cache_mem_size = memory_for_gpu
        self.wrapper = attention_wrapper
        self.curr_seq_lens = [0] * max_batch_size
        self.seq_to_batch_idx: Dict[int, int] = {}
        self.curr_batch_idx: Optional[torch.Tensor] = None
        self.gpu_cache = self.allocate_gpu_cache()

    # :45-79
    def allocate_gpu_cache(self) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        kv = self.alloc.init_kvcache(self.num_layers, self.num_heads, self.head_size, self.max_batch_size,
                                     self.max_model_seq_len, self.device_idx, self.dtype, self.page_size,
                                     self.vattn_mega_cache)
        if self.vattn_mega_cache:
            k, v = kv
            cache = [(k[:, :, i], v[:, :, i]) for i in range(self.num_layers)]  # :58-68
        else:
            cache = list(zip(kv[:self.num_layers], kv[self.num_layers:]))
        self.alloc.reserve_physical_pages(self.cache_mem_size)
        return cache

    def num_free_blocks(self) -> int:  # :42-43
        return self.alloc.num_free_kvblocks()

    def get_k_cache(self, layer_idx: int) -> torch.Tensor:
        return self.gpu_cache[layer_idx][0]

    def get_v_cache(self, layer_idx: int) -> torch.Tensor:
        return self.gpu_cache[layer_idx][1]

    # :131-143
    def get_req_batch_idx(self, seq_id: int, seq_len: int) -> int:
        if seq_id in self.seq_to_batch_idx:
            return self.seq_to_batch_idx[seq_id]
        idx = self.alloc.alloc_new_batch_idx(seq_len)
        assert idx != -1, "Failed to allocate new batch idx. This is not expected..."
        self.seq_to_batch_idx[seq_id] = idx
        return idx

    # :91-124
    def step(self, seq_metadata_list: Sequence) -> None:
        prompt_idx, gen_idx = [], []
        for md in seq_metadata_list:
            if md.is_prompt:
                chunk = md.seq.get_next_prompt_chunk_len(md.prompt_chunk_len)
                context_len = md.seq.get_num_prompt_tokens_processed() + chunk
                target = prompt_idx
            else:
                context_len = md.seq.get_len()
                target = gen_idx
            b = self.get_req_batch_idx(md.seq.seq_id, context_len)
            self.curr_seq_lens[b] = context_len
            target.append(b)
        if self.vattn_async:
            self.alloc.step_async(self.curr_seq_lens)
        else:
            self.alloc.step(self.curr_seq_lens, True)
        self.curr_batch_idx = torch.tensor(prompt_idx + gen_idx, dtype=torch.int32, device=self.device)
        if self.wrapper is not None:
            self.wrapper.set_batch_idx(self.curr_batch_idx,
                                       torch.tensor(gen_idx, dtype=torch.int32, device=self.device))

    # :126-129
    def on_step_completion(self, seq_metadata_list: Sequence) -> None:
        for md in seq_metadata_list:
            if md.seq.is_finished():
                self.free_request(md.seq.seq_id)

    # :81-83
    def preempt_requests(self, preempted_seqs: Sequence) -> None:
        for seq in preempted_seqs:
            self.free_request(seq.seq_id)

    # :145-152
    def free_request(self, seq_id: int) -> None:
        if seq_id not in self.seq_to_batch_idx:
            raise Exception(f"seq_id {seq_id} not found in req_table")
        b = self.seq_to_batch_idx.pop(seq_id)
        self.alloc.free_batch_idx(b)
        self.curr_seq_lens[b] = 0

    def reclaim_req_ids(self) -> None:  # :154-156
        for seq_id in list(self.seq_to_batch_idx):
            self.free_request(seq_id)

    def get_batch_idx(self) -> Optional[torch.Tensor]:
        return self.curr_batch_idx

    def clear_batch_index(self) -> None:
        self.curr_batch_idx = None

    def disable_deferred_reclamation(self) -> None:  # :167-168
        self.alloc.set_deferred_reclamation(False)

    def cleanup_kvcache(self) -> None:  # :192-194
        self.alloc.cleanup()

    @staticmethod
    def get_cache_block_size(block_size: int, num_kv_heads: int, head_size: int, num_layers: int,
                             dtype: torch.dtype) -> int:
        """:172-190: bytes of one scheduler block (K and V, all layers)."""
        itemsize = torch.empty((), dtype=dtype).element_size()
        return itemsize * num_layers * 2 * block_size * num_kv_heads * head_size
