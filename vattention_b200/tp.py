"""Head-sharded tensor parallelism for the attention block.

The reference shards attention by KV/query head: rank r owns q heads [r*Hq/tp, (r+1)*Hq/tp) and
kv heads [r*Hkv/tp, ...) (sarathi/model_executor/models/llama.py:124-131, config.py:139-175), each
worker process has its own `vattention` allocator sized with the per-rank Hkv
(vATTN_cache_engine.py:48-57), page bookkeeping is identical on every rank (same seq lens), and
the only collective on the path is ONE all-reduce(sum) of the row-parallel o_proj output
[num_tokens, hidden] (tensor_parallel/layers.py:448-451 -> mappings.py:16-26).

This module holds that host logic: shard arithmetic + the o_proj-partial/all-reduce step.  The
attention callable is injected (the CUDA operator on GPU; tests inject a CPU function under
gloo), so the same code runs with world_size 2 on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class HeadShard:
    rank: int
    world: int
    num_heads: int       # total q heads
    num_kv_heads: int    # total kv heads
    head_dim: int

    def __post_init__(self):
        # llama.py:129: total_num_kv_heads % tp_size == 0
        if self.num_kv_heads % self.world or self.num_heads % self.world:
            raise ValueError("num_heads and num_kv_heads must be divisible by the TP world size")

    @property
    def heads_per_rank(self) -> int:
        return self.num_heads // self.world

    @property
    def kv_heads_per_rank(self) -> int:
        return self.num_kv_heads // self.world

    @property
    def q_range(self) -> range:
        return range(self.rank * self.heads_per_rank, (self.rank + 1) * self.heads_per_rank)

    @property
    def kv_range(self) -> range:
        return range(self.rank * self.kv_heads_per_rank, (self.rank + 1) * self.kv_heads_per_rank)

    def tokens_per_page(self, page_size: int, itemsize: int = 2, num_layers: int = 1,
                        megacache: bool = False) -> int:
        """engine/arg_utils.py:147-159 with tensor_parallel_size = world: a 2 MB page holds `world`
        times more tokens of a rank's shard."""
        per_token = self.kv_heads_per_rank * self.head_dim * itemsize * (num_layers if megacache else 1)
        return page_size // per_token

    def shard_q(self, q: torch.Tensor) -> torch.Tensor:
        """q [..., Hq, D] -> this rank's heads (a view)."""
        return q[..., self.q_range.start:self.q_range.stop, :]

    def shard_kv(self, kv: torch.Tensor) -> torch.Tensor:
        return kv[..., self.kv_range.start:self.kv_range.stop, :]

    def shard_o_proj(self, w_o: torch.Tensor) -> torch.Tensor:
        """Row-parallel o_proj: W_o [Hq*D, hidden] -> rows of this rank's heads
        (tensor_parallel/layers.py:432-447)."""
        d = self.head_dim
        return w_o[self.q_range.start * d:self.q_range.stop * d]


class HeadShardedAttention:
    """attention(shard) -> partial o_proj -> one all_reduce."""

    def __init__(self, shard: HeadShard, w_o_shard: torch.Tensor,
                 attn_fn: Callable[..., torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        self.shard, self.w_o, self.attn_fn, self.group = shard, w_o_shard, attn_fn, group

    def forward(self, q_shard: torch.Tensor, *attn_args, **attn_kwargs) -> torch.Tensor:
        """q_shard [B, Sq, Hq/tp, D]; returns the all-reduced block output [B*Sq, hidden]."""
        out = self.attn_fn(q_shard, *attn_args, **attn_kwargs)       # [B, Sq, Hq/tp, D]
        flat = out.reshape(out.shape[0] * out.shape[1], -1)           # [tokens, Hq/tp * D]
        partial = flat @ self.w_o                                     # [tokens, hidden]
        if self.shard.world > 1:
            dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.group)
        return partial
