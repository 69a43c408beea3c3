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


class PeerAllReduce:
    """The all-reduce of the row-parallel o_proj output done by our own kernel over NVLink peer
    memory (csrc/tp_allreduce.cu) instead of NCCL: every rank's partial lives in a symmetric
    allocation (torch.distributed._symmetric_memory supplies the peer mappings -- plumbing), the
    o_proj GEMM writes straight into it and one kernel sums all ranks' partials out of peer memory.

    Usage per layer-call:   buf = ar.partial_buffer(tokens)   # view of this call's slot
                            torch.matmul(x, w_o_shard, out=buf)
                            y = ar.reduce(tokens)             # [tokens, hidden], local tensor
    """

    def __init__(self, hidden: int, max_tokens: int, dtype: torch.dtype, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self._C, self._lib = C, _lib
        self.group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("PeerAllReduce supports up to 8 ranks (one NVSwitch domain)")
        self.hidden, self.max_tokens, self.dtype, self.device = hidden, max_tokens, dtype, device
        itemsize = torch.empty((), dtype=dtype).element_size()
        self.slot_bytes = (max_tokens * hidden * itemsize + 255) // 256 * 256
        self.flag_bytes = 256
        total = 2 * self.slot_bytes + 2 * self.flag_bytes
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm_mem.rendezvous(self.buf, self.group.group_name)
        dist.barrier(self.group)          # every rank's flags are zero before anyone publishes
        self.base_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.epoch = 0
        self.out = torch.empty(max_tokens, hidden, dtype=dtype, device=device)
        self._dt = {torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}[dtype]

    def partial_buffer(self, tokens: int) -> torch.Tensor:
        """This call's slot of the symmetric buffer as a [tokens, hidden] tensor (GEMM output)."""
        parity = (self.epoch + 1) & 1
        raw = self.buf[parity * self.slot_bytes:(parity + 1) * self.slot_bytes]
        return raw.view(self.dtype)[: tokens * self.hidden].view(tokens, self.hidden)

    def reduce(self, tokens: int) -> torch.Tensor:
        C, lib = self._C, self._lib.lib
        self.epoch += 1
        parity = self.epoch & 1
        parts = (C.c_uint64 * self.world)(*[b + parity * self.slot_bytes for b in self.base_ptrs])
        flags = (C.c_uint64 * self.world)(*[b + 2 * self.slot_bytes + parity * self.flag_bytes
                                            for b in self.base_ptrs])
        out = self.out[:tokens]
        self._lib.check(lib.vattn_allreduce_oneshot(
            parts, flags, out.data_ptr(), tokens * self.hidden, self._dt, self.rank, self.world,
            self.epoch, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out


class HeadShardedAttentionPeer(HeadShardedAttention):
    """HeadShardedAttention whose collective is PeerAllReduce (GEMM output lands in peer-visible
    memory, one reduction kernel reads every rank's partial over NVLink)."""

    def __init__(self, shard: HeadShard, w_o_shard: torch.Tensor, attn_fn, max_tokens: int,
                 group: Optional[dist.ProcessGroup] = None):
        super().__init__(shard, w_o_shard, attn_fn, group)
        self.ar = PeerAllReduce(w_o_shard.shape[1], max_tokens, w_o_shard.dtype, w_o_shard.device, group)

    def forward(self, q_shard: torch.Tensor, *attn_args, **attn_kwargs) -> torch.Tensor:
        out = self.attn_fn(q_shard, *attn_args, **attn_kwargs)
        flat = out.reshape(out.shape[0] * out.shape[1], -1)
        tokens = flat.shape[0]
        torch.matmul(flat, self.w_o, out=self.ar.partial_buffer(tokens))
        return self.ar.reduce(tokens)


class FusedOProjAllReduce:
    """o_proj GEMM and its all-reduce as ONE kernel per rank (csrc/oproj_allreduce.cu): tcgen05 GEMM
    with the weight rows on the MMA M axis, each finished [tokens x 128] tile pushed straight into
    every rank's receive slot over NVLink, flagged, and summed by the CTA that owns the tile.

    `w_shard` is this rank's slice in nn.Linear layout [hidden, k_local] (RowParallelLinear.weight,
    tensor_parallel/layers.py:432-447).  With `local_only=True` (or no process group) the symmetric
    allocation is a plain local tensor and world = 1: the GEMM path alone, used by single-GPU tests.
    Decode batches only (<= 128 tokens per call); the call is CUDA-graph capturable.
    """

    MAX_TOKENS = 128

    def __init__(self, w_shard: torch.Tensor, max_tokens: int, group: Optional[dist.ProcessGroup] = None,
                 local_only: bool = False):
        import ctypes as C

        from . import _lib
        self._C, self._lib = C, _lib
        if w_shard.dim() != 2 or not w_shard.is_contiguous():
            raise ValueError("w_shard must be a contiguous [hidden, k_local] tensor")
        if max_tokens > self.MAX_TOKENS:
            raise ValueError("FusedOProjAllReduce handles at most 128 tokens per call")
        self.w = w_shard
        self.hidden, self.k_local = w_shard.shape
        self.max_tokens, self.dtype, self.device = max_tokens, w_shard.dtype, w_shard.device
        self._dt = {torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}[self.dtype]
        lib = _lib.lib
        if local_only or not dist.is_initialized():
            self.rank, self.world, self.group = 0, 1, None
        else:
            self.group = group or dist.group.WORLD
            self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.recv_bytes = (lib.vattn_oproj_allreduce_recv_bytes(max_tokens, self.hidden, self.world) + 255) // 256 * 256
        self.flag_bytes = lib.vattn_oproj_allreduce_flag_bytes(self.hidden)
        total = self.recv_bytes + self.flag_bytes
        if self.world == 1:
            self.buf = torch.zeros(total, dtype=torch.uint8, device=self.device)
            bases = [self.buf.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
            self.buf.zero_()
            torch.cuda.synchronize(self.device)
            self.handle = symm_mem.rendezvous(self.buf, self.group.group_name)
            dist.barrier(self.group)      # every rank's flags are zero before anyone publishes
            bases = [int(p) for p in self.handle.buffer_ptrs]
        self._recv = (C.c_uint64 * self.world)(*bases)
        self._flags = (C.c_uint64 * self.world)(*[b + self.recv_bytes for b in bases])
        # [2] error flag, [4 + tile] per-tile epoch (tiles of 32 hidden columns at most)
        self.epoch_state = torch.zeros(4 + self.hidden // 32, dtype=torch.int32, device=self.device)
        self.out = torch.empty(max_tokens, self.hidden, dtype=self.dtype, device=self.device)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x [tokens, k_local] (row stride free) -> all-reduced [tokens, hidden] (a view of an
        internal buffer, valid until the next call)."""
        tokens = x.shape[0]
        if x.dim() != 2 or x.shape[1] != self.k_local or x.stride(1) != 1 or x.dtype != self.dtype:
            raise ValueError("x must be [tokens, k_local] with unit inner stride and the weight's dtype")
        out = self.out[:tokens]
        C = self._C
        self._lib.check(self._lib.lib.vattn_oproj_allreduce(
            x.data_ptr(), x.stride(0), self.w.data_ptr(), out.data_ptr(), tokens, self.hidden, self.k_local,
            self._dt, self.max_tokens, self._recv, self._flags, self.epoch_state.data_ptr(), self.rank,
            self.world, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out

    def failed(self) -> bool:
        """True after a call gave up waiting for a peer (device-side spin limit); synchronises."""
        return bool(self.epoch_state[2].item())


class HeadShardedAttentionFused(HeadShardedAttention):
    """HeadShardedAttention whose o_proj + all-reduce is the fused kernel."""

    def __init__(self, shard: HeadShard, w_o_shard: torch.Tensor, attn_fn, max_tokens: int,
                 group: Optional[dist.ProcessGroup] = None):
        super().__init__(shard, w_o_shard, attn_fn, group)
        # HeadShard.shard_o_proj hands out [k_local, hidden]; the kernel streams nn.Linear's [hidden, k_local]
        self.op = FusedOProjAllReduce(w_o_shard.t().contiguous(), max_tokens, group)

    def forward(self, q_shard: torch.Tensor, *attn_args, **attn_kwargs) -> torch.Tensor:
        out = self.attn_fn(q_shard, *attn_args, **attn_kwargs)
        return self.op(out.reshape(out.shape[0] * out.shape[1], -1))
