"""Host-side mirrors of the sarathi attention wrappers that sit on the vAttention hot path.

Reference classes (sarathi-lean/sarathi/model_executor/attention/):
    VAttentionFlashAttentionWrapper      vattention_flashattention_wrapper.py:17-224   (fa_vattn)
    VAttentionFlashInferWrapper          vattention_flashinfer_wrapper.py              (fi_vattn)
    VAttentionFlashAttentionPODWrapper   vattention_flashattention_pod_wrapper.py      (fa_pod)
and their interface, base_attention_wrapper.py:11-68: init / begin_forward(seq_metadata_list) /
forward(query, key, value, kv_cache, softmax_scale, layer_id) / end_forward / set_batch_idx.

Same method names, argument meaning and batch layout (prefill tokens first, then one token per
decode sequence; batch_idx lists prefills then decodes, vATTN_cache_engine.py:123-124).  The
sequence metadata is duck-typed: anything with .is_prompt, .prompt_chunk_len and .seq exposing
get_next_prompt_chunk_len(n), get_num_prompt_tokens_processed(), get_len() works, so sarathi's
SequenceMetadata plugs in unchanged.  The operators come from vattention_b200.attention (the C
ABI); a different namespace can be injected for host-logic tests.

Deliberate differences from the reference, all observable only as fewer launches / fixed bugs:
  * prefill chunks are not looped through five CudaTimer sections; timers are optional no-ops;
  * the POD wrapper writes the chunk's K/V at row `processed` (the reference passes the cache
    un-offset to cache_flat, vattention_flashattention_pod_wrapper.py:153-158, which is only right
    for the first chunk) and forwards softmax_scale (the reference drops it);
  * metadata lists are rebuilt in begin_forward instead of appended to an instance list that
    end_forward must clear (vattention_flashattention_wrapper.py:67,96).
"""
from __future__ import annotations

from contextlib import nullcontext
from typing import List, Optional, Sequence, Tuple

import torch

from . import attention as _default_ops


class _VAttentionWrapperBase:
    _inst = None

    @classmethod
    def get_instance(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self, ops=None):
        self.ops = ops or _default_ops
        self.is_metadata_initialized = False
        self.is_profiling_iteration = False

    # base_attention_wrapper.py:14-28
    def init(self, model_config=None, parallel_config=None, block_size: int = 0,
             device: Optional[torch.device] = None, *, num_q_heads: int = 0, num_kv_heads: int = 0,
             head_dim: int = 0):
        if model_config is not None:
            num_q_heads = model_config.get_num_q_heads(parallel_config)
            num_kv_heads = model_config.get_num_kv_heads(parallel_config)
            head_dim = model_config.get_head_size()
        self.device = device
        self.num_q_heads, self.num_kv_heads, self.head_dim = num_q_heads, num_kv_heads, head_dim
        self.block_size = block_size
        self._reset()
        return self

    def _reset(self):
        self.is_metadata_initialized = False
        self.prefill_query_lens: List[int] = []
        self.prefill_cache_lens: List[int] = []
        self.current_total_len_device_lst: List[torch.Tensor] = []
        self.decode_cache_lens: Optional[torch.Tensor] = None
        self.batch_index: Optional[torch.Tensor] = None
        self.batch_index_gen: Optional[torch.Tensor] = None
        self.max_cache_len = 0
        self.decode_batch_size = 0

    def get_timer(self, operation=None, layer_id=None):
        return nullcontext()   # sarathi's CudaTimer is observability, outside the hot path

    # vattention_flashattention_wrapper.py:44-90
    def begin_forward(self, seq_metadata_list: Sequence) -> None:
        self.is_profiling_iteration = False
        q_lens, cached, totals, decode_lens = [], [], [], []
        for md in seq_metadata_list:
            if md.is_prompt:
                chunk = md.seq.get_next_prompt_chunk_len(md.prompt_chunk_len)
                done = md.seq.get_num_prompt_tokens_processed()
                q_lens.append(chunk)
                cached.append(done)
                totals.append(done + chunk)
        for md in seq_metadata_list:
            if not md.is_prompt:
                decode_lens.append(md.seq.get_len() - 1)
        self.prefill_query_lens, self.prefill_cache_lens = q_lens, cached
        self.current_total_len_device_lst = [
            torch.tensor([t], dtype=torch.int32, device=self.device) for t in totals]
        self.decode_batch_size = len(decode_lens)
        if decode_lens:
            self.decode_cache_lens = torch.tensor(decode_lens, dtype=torch.int32, device=self.device)
            self.max_cache_len = max(decode_lens) + 1
        else:
            self.decode_cache_lens, self.max_cache_len = None, 0
        self.is_metadata_initialized = True

    def end_forward(self):  # :92-104
        # batch indices are set by the cache engine BEFORE begin_forward of the same step
        # (base_worker.py:186-188), so begin_forward must not touch them; they are dropped here
        self._reset()

    def set_batch_idx(self, batch_idx: torch.Tensor, batch_idx_gen: torch.Tensor) -> None:  # :106-108
        self.batch_index = batch_idx.to(torch.int32)
        self.batch_index_gen = batch_idx_gen.to(torch.int32)

    # -- pieces shared by the three backends ------------------------------------------------
    def _views(self, t: torch.Tensor, n_heads: int) -> torch.Tensor:
        return t.reshape(-1, n_heads, self.head_dim)

    def _save_chunk(self, key, value, kv_cache, index: int, start: int, n: int, offset: int,
                    end: Optional[int] = None):
        k = self._views(key[start:start + n], self.num_kv_heads)
        v = self._views(value[start:start + n], self.num_kv_heads)
        self.ops.cache_flat(k, v, kv_cache[0][index][offset:end], kv_cache[1][index][offset:end], "auto")

    def _decode(self, query, key, value, kv_cache, token_offset: int, softmax_scale, full_cache: bool):
        n = self.decode_batch_size
        q = query[token_offset:token_offset + n].reshape(n, 1, self.num_q_heads, self.head_dim)
        k = key[token_offset:token_offset + n].reshape(n, 1, self.num_kv_heads, self.head_dim)
        v = value[token_offset:token_offset + n].reshape(n, 1, self.num_kv_heads, self.head_dim)
        kc = kv_cache[0] if full_cache else kv_cache[0][:, :self.max_cache_len]
        vc = kv_cache[1] if full_cache else kv_cache[1][:, :self.max_cache_len]
        return self.ops.flash_attn_with_kvcache(
            q, kc, vc, k, v, cache_seqlens=self.decode_cache_lens, block_table=None,
            softmax_scale=softmax_scale, causal=True, cache_batch_idx=self.batch_index_gen)

    def _check(self, query):
        assert self.is_metadata_initialized, "Metadata is not initialized."
        if self.is_profiling_iteration:
            return torch.zeros_like(query)  # :121-123
        return None


class VAttentionFlashAttentionWrapper(_VAttentionWrapperBase):
    """fa_vattn: per prefill chunk cache_flat + flash_attn_with_kvcache(causal, cache_seqlens=[total]);
    all decodes in one flash_attn_with_kvcache with append (wrapper.py:110-224)."""

    def forward(self, query, key, value, kv_cache: Tuple[torch.Tensor, torch.Tensor],
                softmax_scale: float = 1.0, layer_id: Optional[int] = None) -> torch.Tensor:
        z = self._check(query)
        if z is not None:
            return z
        output = torch.empty_like(query)
        off = 0
        for idx, (done, n, total) in enumerate(zip(self.prefill_cache_lens, self.prefill_query_lens,
                                                   self.current_total_len_device_lst)):
            index = int(self.batch_index[idx])
            self._save_chunk(key, value, kv_cache, index, off, n, done)
            q = query[off:off + n].reshape(1, n, self.num_q_heads, self.head_dim)
            kc = kv_cache[0][index].unsqueeze(0)
            vc = kv_cache[1][index].unsqueeze(0)
            o = self.ops.flash_attn_with_kvcache(q, kc, vc, cache_seqlens=total, causal=True,
                                                 softmax_scale=softmax_scale)
            output[off:off + n] = o.reshape(n, -1)
            off += n
        if self.decode_batch_size:
            o = self._decode(query, key, value, kv_cache, off, softmax_scale, full_cache=False)
            output[off:off + self.decode_batch_size] = o.reshape(self.decode_batch_size, -1)
        return output


class VAttentionFlashInferWrapper(_VAttentionWrapperBase):
    """fi_vattn: prefill through single_prefill_with_kv_cache on the [:processed+chunk] slice with the
    default 1/sqrt(D) scale (vattention_flashinfer_wrapper.py:151-158); decode as fa_vattn but over
    the whole cache tensor (:188-199)."""

    def forward(self, query, key, value, kv_cache, softmax_scale: float = 1.0,
                layer_id: Optional[int] = None) -> torch.Tensor:
        z = self._check(query)
        if z is not None:
            return z
        output = torch.empty_like(query)
        off = 0
        for idx, (done, n) in enumerate(zip(self.prefill_cache_lens, self.prefill_query_lens)):
            index = int(self.batch_index[idx])
            # the FI wrapper slices the cache to the chunk's end first (flashinfer_wrapper.py:140-148)
            self._save_chunk(key, value, kv_cache, index, off, n, done, done + n)
            q = self._views(query[off:off + n], self.num_q_heads)
            o = self.ops.single_prefill_with_kv_cache(q, kv_cache[0][index][:done + n],
                                                      kv_cache[1][index][:done + n], causal=True)
            output[off:off + n] = o.reshape(n, -1)
            off += n
        if self.decode_batch_size:
            o = self._decode(query, key, value, kv_cache, off, softmax_scale, full_cache=True)
            output[off:off + self.decode_batch_size] = o.reshape(self.decode_batch_size, -1)
        return output


class VAttentionFlashAttentionPODWrapper(_VAttentionWrapperBase):
    """fa_pod: at most one prefill chunk fused with the decode batch in one
    true_fused_attn_with_kvcache call (pod_wrapper.py:121-203)."""

    def __init__(self, ops=None):
        super().__init__(ops)
        self.fused_param = 11

    def begin_forward(self, seq_metadata_list: Sequence) -> None:
        super().begin_forward(seq_metadata_list)
        if len(self.prefill_query_lens) > 1:
            raise ValueError("Batched prefills are not supported currently ...")  # :74-75
        # :98-101: smaller prefill tiles once the processed prompt is long
        long_prompt = bool(self.prefill_cache_lens) and self.prefill_cache_lens[0] > 10240
        self.fused_param = 11 if (self.decode_batch_size and long_prompt) else 9

    def forward(self, query, key, value, kv_cache, softmax_scale: float = 1.0,
                layer_id: Optional[int] = None) -> torch.Tensor:
        z = self._check(query)
        if z is not None:
            return z
        output = torch.empty_like(query)
        q_p = kc_p = vc_p = total = None
        n_p = 0
        if self.prefill_query_lens:
            n_p, done = self.prefill_query_lens[0], self.prefill_cache_lens[0]
            index = int(self.batch_index[0])
            self._save_chunk(key, value, kv_cache, index, 0, n_p, done)
            q_p = query[:n_p].reshape(1, n_p, self.num_q_heads, self.head_dim)
            kc_p, vc_p = kv_cache[0][index].unsqueeze(0), kv_cache[1][index].unsqueeze(0)
            total = self.current_total_len_device_lst[0]
        q_d = k_d = v_d = None
        n_d = self.decode_batch_size
        if n_d:
            q_d = query[n_p:n_p + n_d].reshape(n_d, 1, self.num_q_heads, self.head_dim)
            k_d = key[n_p:n_p + n_d].reshape(n_d, 1, self.num_kv_heads, self.head_dim)
            v_d = value[n_p:n_p + n_d].reshape(n_d, 1, self.num_kv_heads, self.head_dim)
        if q_p is None and q_d is None:
            return output
        if q_p is None:
            out_p, out_d = self.ops.true_fused_attn_with_kvcache(
                None, None, None, q_d, kv_cache[0], kv_cache[1], k_d, v_d, causal=True,
                cache_seqlens_d=self.decode_cache_lens, cache_batch_idx=self.batch_index_gen,
                softmax_scale=softmax_scale, fused_params=self.fused_param)
        else:
            out_p, out_d = self.ops.true_fused_attn_with_kvcache(
                q_p, kc_p, vc_p, q_d, kv_cache[0] if n_d else None, kv_cache[1] if n_d else None,
                k_d, v_d, causal=True, cache_seqlens_p=total, cache_seqlens_d=self.decode_cache_lens,
                cache_batch_idx=self.batch_index_gen, softmax_scale=softmax_scale,
                fused_params=self.fused_param)
        if out_p is not None:
            output[:n_p] = out_p.reshape(n_p, -1)
        if out_d is not None:
            output[n_p:n_p + n_d] = out_d.reshape(n_d, -1)
        return output


_BACKENDS = {
    # sarathi/model_executor/attention/__init__.py:36-54,124-160 backend names
    "fa_vattn": VAttentionFlashAttentionWrapper, "fa_vattn_sync": VAttentionFlashAttentionWrapper,
    "fa_vattn_megacache": VAttentionFlashAttentionWrapper,
    "fa_vattn_megacache_sync": VAttentionFlashAttentionWrapper,
    "fi_vattn": VAttentionFlashInferWrapper, "fi_vattn_sync": VAttentionFlashInferWrapper,
    "fa_pod": VAttentionFlashAttentionPODWrapper, "fa_pod_megacache": VAttentionFlashAttentionPODWrapper,
}


def get_attention_wrapper_class(backend: str):
    try:
        return _BACKENDS[backend.lower()]
    except KeyError:
        raise ValueError(f"Unsupported attention backend on the vAttention path: {backend}") from None
