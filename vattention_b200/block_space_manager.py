"""Scheduler-facing free-block accounting for the vAttention allocator.

Mirror of sarathi-lean/sarathi/core/block_space_manager/vattention_block_space_manager.py:9-98:
the scheduler never owns block tables with vAttention; each iteration the engine pushes the
allocator's `num_free_kvblocks()` into `set_free_blocks`, and admission control works on
    free_blocks - promised_blocks - need >= watermark_blocks
where `promised_blocks` counts blocks the scheduler has handed out since the last refresh
(`clear_promised_blocks`) and that the allocator will only map in its next `step_async`.

A sequence is any object with `seq_id` and `get_len()` (sarathi's Sequence, :31-35).
Same method names, arguments and return values as the reference class; `refresh()` is the one
addition: it reads the allocator directly.
"""
from __future__ import annotations

from typing import Dict, List, Optional


def _blocks(num_tokens: int, block_size: int) -> int:
    return -(-num_tokens // block_size)       # ceil, as math.ceil(len / block_size) in :31-35


class vAttentionBlockSpaceManager:
    def __init__(self, block_size: int, num_gpu_blocks: int, max_model_len: int,
                 watermark: float = 0.01) -> None:
        if watermark < 0.0:
            raise AssertionError("watermark must be non-negative")        # :22
        self.block_size = block_size
        self.num_total_gpu_blocks = num_gpu_blocks
        self.max_model_len = max_model_len
        self.watermark = watermark
        self.watermark_blocks = int(watermark * num_gpu_blocks)
        self.promised_blocks = 0
        self.active_requests: Dict[int, object] = {}
        self.preemption_queue: List[object] = []
        # the reference leaves `free_blocks` undefined until the first set_free_blocks (:45-46);
        # reading it earlier is an AttributeError there and here

    # ---- refreshed every engine iteration -------------------------------------------------
    def set_free_blocks(self, free_blocks: int) -> None:
        self.free_blocks = free_blocks

    def clear_promised_blocks(self) -> None:
        self.promised_blocks = 0

    def refresh(self, allocator=None) -> int:
        """set_free_blocks(allocator.num_free_kvblocks()) + clear_promised_blocks(): what the
        engine does at the top of a scheduling round (base_scheduler usage of :45-46, :87-88)."""
        if allocator is None:
            from . import vattention as allocator
        self.set_free_blocks(int(allocator.num_free_kvblocks()))
        self.clear_promised_blocks()
        return self.free_blocks

    # ---- admission ------------------------------------------------------------------------
    def get_num_blocks(self, seq) -> int:
        return _blocks(seq.get_len(), self.block_size)

    def can_allocate(self, seq) -> bool:
        room = self.free_blocks - self.promised_blocks - self.get_num_blocks(seq)
        return room >= self.watermark_blocks                               # :37-43

    def allocate(self, seq) -> None:
        self.active_requests[seq.seq_id] = seq
        self.promised_blocks += self.get_num_blocks(seq)                   # :48-50

    def can_append_slot(self) -> bool:
        return self.free_blocks - self.promised_blocks > 0                 # :52-57

    def append_slot(self, seq) -> None:
        """A decode token that opens a new block is one more promised block (:60-67)."""
        n = seq.get_len()
        if _blocks(n + 1, self.block_size) > _blocks(n, self.block_size):
            self.promised_blocks += 1

    def free(self, seq) -> None:
        if self.active_requests.pop(seq.seq_id, None) is not None:         # :75-81
            self.free_blocks += self.get_num_blocks(seq)

    def reset(self) -> None:
        self.active_requests = {}

    def is_allocated(self, seq) -> bool:
        return seq.seq_id in self.active_requests

    def get_num_free_gpu_blocks(self, seq=None) -> int:
        return self.free_blocks                                            # :96-97

    # block tables do not exist with a contiguous virtual KV cache (:69-73, :90-91)
    def get_block_table(self, seq) -> Optional[List[int]]:
        return None

    def _get_physical_blocks(self, seq) -> None:
        return None

    def _free_block_table(self, block_table) -> None:
        return None
