"""Operator boundary: the callables the sarathi vattention wrappers dispatch to.

    flash_attn_with_kvcache       vattention_flashattention_wrapper.py:159-166,194-205
                                  (FA API: flash_attn/flash_attn_interface.py; arithmetic
                                   pod_attn/pod_attn/flash_api.cpp:1291-1580)
    single_prefill_with_kv_cache  vattention_flashinfer_wrapper.py:151-158
    true_fused_attn_with_kvcache  vattention_flashattention_pod_wrapper.py:177-191,
                                  pod_attn/pod_attn/fused_attn_interface.py:12-137
    cache_flat                    sarathi-lean/csrc/cache.cpp:40-46, cache_kernels.cu:524-570

Same keyword surfaces, tensors in / tensors out on the current CUDA stream, errors as
RuntimeError.  Everything below is argument marshalling into the C ABI
(include/vattn_b200.h); the arithmetic is in libvattn_b200.so and nowhere else.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import torch

from . import _lib
from ._lib import FwdParams, check, lib

_DT = {torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}
_IMPL = {"auto": _lib.IMPL_AUTO, "simt": _lib.IMPL_SIMT, "tc": _lib.IMPL_TC}

# one scratch buffer per device, grown on demand (split-KV partials); the reference's FA
# allocates softmax_lse_accum / out_accum per call (flash_api.cpp:300-323)
_workspace: dict = {}


def _ws(device: torch.device, nbytes: int) -> Optional[torch.Tensor]:
    if nbytes == 0:
        return None
    key = (device, torch.cuda.current_stream(device).cuda_stream)  # calls on two streams may overlap
    buf = _workspace.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspace[key] = buf
    return buf


def _stream(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be on CUDA device")  # FA's CHECK_DEVICE wording
    if t.stride(-1) != 1:
        raise RuntimeError(f"{name} must have contiguous last dimension")


def _fill_params(q, k_cache, v_cache, k, v, out, cache_seqlens, cache_batch_idx,
                 softmax_scale, causal, impl, num_splits, lse, rotary=None) -> FwdParams:
    p = FwdParams()
    b, sq, hq, d = q.shape
    cb, sk, hkv, dk = k_cache.shape
    if dk != d or v_cache.shape != k_cache.shape:
        raise RuntimeError("k_cache / v_cache shape mismatch")
    p.q = q.data_ptr()
    p.q_batch_stride, p.q_row_stride, p.q_head_stride = q.stride(0), q.stride(1), q.stride(2)
    p.k_cache, p.v_cache = k_cache.data_ptr(), v_cache.data_ptr()
    p.k_batch_stride, p.k_row_stride, p.k_head_stride = k_cache.stride(0), k_cache.stride(1), k_cache.stride(2)
    p.v_batch_stride, p.v_row_stride, p.v_head_stride = v_cache.stride(0), v_cache.stride(1), v_cache.stride(2)
    if k is not None:
        p.k_new, p.v_new = k.data_ptr(), v.data_ptr()
        p.knew_batch_stride, p.knew_row_stride, p.knew_head_stride = k.stride(0), k.stride(1), k.stride(2)
        p.vnew_batch_stride, p.vnew_row_stride, p.vnew_head_stride = v.stride(0), v.stride(1), v.stride(2)
        p.seqlen_new = k.shape[1]
    p.out = out.data_ptr()
    p.o_batch_stride, p.o_row_stride, p.o_head_stride = out.stride(0), out.stride(1), out.stride(2)
    p.softmax_lse = lse.data_ptr() if lse is not None else None
    p.cache_seqlens = cache_seqlens.data_ptr() if cache_seqlens is not None else None
    p.cache_batch_idx = cache_batch_idx.data_ptr() if cache_batch_idx is not None else None
    p.batch, p.cache_batch, p.seqlen_q, p.seqlen_k = b, cb, sq, sk
    p.num_heads, p.num_kv_heads, p.head_dim = hq, hkv, d
    p.dtype = _DT[q.dtype]
    p.causal = 1 if causal else 0
    p.softmax_scale = float(softmax_scale)
    p.impl = _IMPL[impl] if isinstance(impl, str) else int(impl)
    p.num_splits = int(num_splits)
    if rotary is not None:
        cos, sin, interleaved = rotary
        p.rotary_cos, p.rotary_sin = cos.data_ptr(), sin.data_ptr()
        p.rotary_dim, p.seqlen_ro = 2 * cos.shape[1], cos.shape[0]
        p.rotary_interleaved = 1 if interleaved else 0
    return p


def _check_rotary(q, k, rotary_cos, rotary_sin):
    """Argument rules of flash_api.cpp:1503-1527, same messages."""
    if rotary_cos is None:
        if rotary_sin is not None:
            raise RuntimeError("If rotary sin is provided, rotary cos must also be provided")
        return None
    if k is None:
        raise RuntimeError("If rotary cos/sin are provided, new key / value to be appended to KV cache "
                           "must also be provided")
    if rotary_sin is None:
        raise RuntimeError("If rotary cos is provided, rotary sin must also be provided")
    for name, t in (("rotary_cos", rotary_cos), ("rotary_sin", rotary_sin)):
        _require_cuda(t, name)
        if t.dim() != 2 or not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous [seqlen_ro, rotary_dim / 2]")
        if t.dtype != q.dtype:
            raise RuntimeError("rotary_cos must have the same dtype as query")
    if rotary_sin.shape != rotary_cos.shape:
        raise RuntimeError("rotary_sin must have the shape of rotary_cos")
    return rotary_cos, rotary_sin


def _prep(q, k_cache, v_cache, k, v, cache_seqlens, cache_batch_idx, softmax_scale):
    _require_cuda(q, "q")
    _require_cuda(k_cache, "k_cache")
    _require_cuda(v_cache, "v_cache")
    if q.dtype not in _DT:
        raise RuntimeError("FlashAttention only support fp16 and bf16 data type")
    if k_cache.dtype != q.dtype or v_cache.dtype != q.dtype:
        raise RuntimeError("query and key must have the same dtype")
    if q.dim() != 4 or k_cache.dim() != 4:
        raise RuntimeError("q and k_cache must be 4-D [batch, seqlen, heads, head_dim]")
    if (k is None) != (v is None):
        raise RuntimeError("k and v must be supplied together")
    if k is not None:
        _require_cuda(k, "k")
        _require_cuda(v, "v")
        if k.shape[0] != q.shape[0]:
            raise RuntimeError("k must have the same batch size as q")
    if softmax_scale is None:
        softmax_scale = q.shape[-1] ** (-0.5)
    if cache_seqlens is not None:
        if isinstance(cache_seqlens, int):
            cache_seqlens = torch.full((q.shape[0],), cache_seqlens, dtype=torch.int32,
                                       device=q.device)
        if cache_seqlens.dtype != torch.int32:
            raise RuntimeError("cache_seqlens must have dtype int32")
        cache_seqlens = cache_seqlens.contiguous()
    if cache_batch_idx is not None:
        if cache_batch_idx.dtype != torch.int32:
            raise RuntimeError("cache_batch_idx must have dtype int32")
        cache_batch_idx = cache_batch_idx.contiguous()
    elif k_cache.shape[0] != q.shape[0]:
        raise RuntimeError("batch size of q and k_cache differ and no cache_batch_idx given")
    return cache_seqlens, cache_batch_idx, softmax_scale


def flash_attn_with_kvcache(
    q, k_cache, v_cache, k=None, v=None, rotary_cos=None, rotary_sin=None,
    cache_seqlens: Optional[Union[int, torch.Tensor]] = None,
    cache_batch_idx: Optional[torch.Tensor] = None, cache_leftpad=None,
    block_table: Optional[torch.Tensor] = None, softmax_scale=None, causal=False,
    window_size=(-1, -1), softcap=0.0, rotary_interleaved=True, alibi_slopes=None,
    num_splits=0, return_softmax_lse=False, *, impl: str = "auto", out=None,
):
    """Attention of q against a contiguous KV cache, optionally appending (k, v) first.

    q [B, Sq, Hq, D]; k_cache/v_cache [Bc, Sk, Hkv, D] (any outer strides, e.g. vAttention
    virtual tensors or megacache views); k/v [B, Snew, Hkv, D] are written in place at rows
    cache_seqlens[b].. of slot cache_batch_idx[b] before attending; returns [B, Sq, Hq, D].
    rotary_cos / rotary_sin [seqlen_ro, rotary_dim/2] rotate q and the appended k first (new key t
    at position cache_seqlens[b]+t; query i at cache_seqlens[b]+i when causal, cache_seqlens[b]
    otherwise; rotary_interleaved pairs dims (2j,2j+1), else (j, j+rotary_dim/2)).
    Options the sarathi wrappers never pass (block_table, ALiBi, sliding window, softcap, left
    padding) are rejected rather than ignored.
    """
    if block_table is not None:
        raise RuntimeError("block_table is not supported: vAttention K/V is contiguous by construction")
    if alibi_slopes is not None or cache_leftpad is not None or softcap != 0.0 \
            or tuple(window_size) != (-1, -1):
        raise RuntimeError("alibi / leftpad / softcap / sliding window are outside the vAttention hot path")
    cache_seqlens, cache_batch_idx, softmax_scale = _prep(
        q, k_cache, v_cache, k, v, cache_seqlens, cache_batch_idx, softmax_scale)
    rot = _check_rotary(q, k, rotary_cos, rotary_sin)
    if rot is not None:
        rot = (rot[0], rot[1], rotary_interleaved)
    if out is None:
        out = torch.empty_like(q)
    if q.shape[0] == 0:
        return (out, None) if return_softmax_lse else out
    lse = None
    if return_softmax_lse:
        lse = torch.empty((q.shape[0], q.shape[2], q.shape[1]), dtype=torch.float32, device=q.device)
    p = _fill_params(q, k_cache, v_cache, k, v, out, cache_seqlens, cache_batch_idx,
                     softmax_scale, causal, impl, num_splits, lse, rot)
    need = lib.vattn_fwd_kvcache_workspace(C.byref(p))
    ws = _ws(q.device, need)
    if ws is not None:
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib.vattn_fwd_kvcache(C.byref(p), _stream(q.device)))
    return (out, lse) if return_softmax_lse else out


def single_prefill_with_kv_cache(q, k, v, causal: bool = False, kv_layout: str = "NHD",
                                 pos_encoding_mode: str = "NONE", sm_scale: Optional[float] = None,
                                 *, impl: str = "auto", **unsupported):
    """flashinfer.single_prefill_with_kv_cache for the call the FI wrapper makes
    (vattention_flashinfer_wrapper.py:151-158): q [c, Hq, D], k/v [n, Hkv, D], NHD layout,
    bottom-right aligned causal mask, scale 1/sqrt(D) unless sm_scale is given."""
    if kv_layout != "NHD" or pos_encoding_mode != "NONE" or unsupported:
        raise RuntimeError("only kv_layout='NHD', pos_encoding_mode='NONE' are on the vAttention path")
    out = flash_attn_with_kvcache(
        q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), causal=causal,
        softmax_scale=sm_scale, impl=impl)
    return out.squeeze(0)


def true_fused_attn_with_kvcache(
    q_p, k_cache_p, v_cache_p, q_d, k_cache_d, v_cache_d, k=None, v=None,
    rotary_cos=None, rotary_sin=None,
    cache_seqlens_p: Optional[Union[int, torch.Tensor]] = None,
    cache_seqlens_d: Optional[Union[int, torch.Tensor]] = None,
    cache_batch_idx: Optional[torch.Tensor] = None, cache_leftpad=None,
    block_table_p=None, block_table_d=None, softmax_scale=None, causal=False,
    window_size=(-1, -1), softcap=0.0, rotary_interleaved=True, alibi_slopes=None,
    num_splits_p=0, num_splits_d=0, return_softmax_lse=False, fused_params=15,
    *, impl: str = "auto",
) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """POD: prefill attention (q_p over k_cache_p) and decode attention (q_d over k_cache_d with
    optional append of k/v and cache_batch_idx) in one call; returns (out_p, out_d).  Same
    surface as fused_attn_interface.py:12-39.  k/v/cache_batch_idx belong to the decode side
    only (:19-20).  If one side is None the other runs alone (:40-78)."""
    if block_table_p is not None or block_table_d is not None:
        raise RuntimeError("block tables are not supported: vAttention K/V is contiguous")
    if rotary_cos is not None or alibi_slopes is not None or cache_leftpad is not None \
            or softcap != 0.0 or tuple(window_size) != (-1, -1):
        raise RuntimeError("rotary / alibi / leftpad / softcap / sliding window are outside the vAttention hot path")
    if q_p is None and q_d is None:
        return None, None
    if q_p is None:
        return None, flash_attn_with_kvcache(
            q_d, k_cache_d, v_cache_d, k=k, v=v, cache_seqlens=cache_seqlens_d,
            cache_batch_idx=cache_batch_idx, softmax_scale=softmax_scale, causal=causal,
            num_splits=num_splits_d, impl=impl)
    if q_d is None:
        # the reference forwards k/v/cache_batch_idx to the prefill call here (:60-77)
        return flash_attn_with_kvcache(
            q_p, k_cache_p, v_cache_p, k=k, v=v, cache_seqlens=cache_seqlens_p,
            cache_batch_idx=cache_batch_idx, softmax_scale=softmax_scale, causal=causal,
            num_splits=num_splits_p, impl=impl), None

    if softmax_scale is None:
        softmax_scale = q_p.shape[-1] ** (-0.5)
    sl_p, _, _ = _prep(q_p, k_cache_p, v_cache_p, None, None, cache_seqlens_p, None, softmax_scale)
    sl_d, bi_d, _ = _prep(q_d, k_cache_d, v_cache_d, k, v, cache_seqlens_d, cache_batch_idx, softmax_scale)
    out_p, out_d = torch.empty_like(q_p), torch.empty_like(q_d)
    pp = _fill_params(q_p, k_cache_p, v_cache_p, None, None, out_p, sl_p, None,
                      softmax_scale, causal, impl, num_splits_p, None)
    pd = _fill_params(q_d, k_cache_d, v_cache_d, k, v, out_d, sl_d, bi_d,
                      softmax_scale, causal, impl, num_splits_d, None)
    need = lib.vattn_pod_workspace(C.byref(pp), C.byref(pd))
    ws = _ws(q_p.device, need)
    check(lib.vattn_pod_fwd(C.byref(pp), C.byref(pd), int(fused_params),
                            C.c_void_p(ws.data_ptr() if ws is not None else 0),
                            ws.numel() if ws is not None else 0, _stream(q_p.device)))
    return out_p, out_d


def cache_flat(key: torch.Tensor, value: torch.Tensor, k_cache: torch.Tensor,
               v_cache: torch.Tensor, kv_cache_dtype: str = "auto") -> None:
    """k_cache[t] = key[t], v_cache[t] = value[t] for t < key.size(0); key/value [c, Hkv, D]
    (cache_kernels.cu:524-570; called at vattention_flashattention_wrapper.py:151-155 with
    k_cache already sliced to the destination rows)."""
    if kv_cache_dtype != "auto":
        raise RuntimeError(f"Unsupported data type of kv cache: {kv_cache_dtype}")  # :531-533
    if k_cache.stride(0) != v_cache.stride(0):
        raise RuntimeError("Expected k_cache.stride(0) == v_cache.stride(0) to be true, but got false.")  # :542
    for t, n in ((key, "key"), (value, "value"), (k_cache, "k_cache"), (v_cache, "v_cache")):
        if not t.is_cuda:
            raise RuntimeError(f"{n} must be a CUDA tensor")
        # the reference indexes i < H*D off each row base (:505-518): rows must be dense
        if t.dim() != 3 or t.stride(2) != 1 or t.stride(1) != t.shape[2]:
            raise RuntimeError(f"{n} must be [tokens, heads, head_size] with dense rows")
    if key.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        raise RuntimeError("cache_flat: unsupported dtype")
    c, h, d = key.shape
    if c == 0:
        return
    if k_cache.shape[0] < c or v_cache.shape[0] < c:
        raise RuntimeError("k_cache has fewer rows than key")
    check(lib.vattn_cache_flat(key.data_ptr(), value.data_ptr(), k_cache.data_ptr(),
                               v_cache.data_ptr(), c, h * d, key.stride(0), value.stride(0),
                               k_cache.stride(0), v_cache.stride(0), key.element_size(),
                               _stream(key.device)))


def launch_count() -> int:
    return int(lib.vattn_launch_count())


def flash_attn_with_kvcache_host(q_host: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                                 k_host: Optional[torch.Tensor], v_host: Optional[torch.Tensor],
                                 cache_seqlens_host: Optional[torch.Tensor],
                                 cache_batch_idx_host: Optional[torch.Tensor], out_host: torch.Tensor,
                                 softmax_scale: Optional[float] = None, causal: bool = False,
                                 impl: str = "auto", wait: bool = True,
                                 pipelined: bool = False) -> torch.Tensor:
    """Same operation with HOST (ideally pinned) q / k / v / index / out buffers and device-resident
    caches: the library copies in, runs the kernels and copies the result back, then drains the
    stream (vattn_fwd_kvcache_host).  wait=False enqueues only (vattn_fwd_kvcache_host_async): the
    result is valid after the stream is synchronised, and the layers of one decode iteration can be
    issued back to back.  This is the call bench.py's `e2e` leg times.
    pipelined=True (with wait=False) moves the copies to their own streams so they overlap the
    neighbouring calls' kernels (vattn_fwd_kvcache_host_pipelined); call host_pipeline_join() before
    synchronising the stream."""
    for t in (q_host, k_host, v_host, cache_seqlens_host, cache_batch_idx_host, out_host):
        if t is not None and (t.is_cuda or not t.is_contiguous()):
            raise RuntimeError("host tensors must be contiguous CPU tensors")
    if softmax_scale is None:
        softmax_scale = q_host.shape[-1] ** (-0.5)
    p = _fill_params(q_host, k_cache, v_cache, k_host, v_host, out_host, cache_seqlens_host,
                     cache_batch_idx_host, softmax_scale, causal, impl, 0, None)
    if pipelined and wait:
        raise RuntimeError("pipelined=True needs wait=False (join with host_pipeline_join())")
    fn = lib.vattn_fwd_kvcache_host if wait else (
        lib.vattn_fwd_kvcache_host_pipelined if pipelined else lib.vattn_fwd_kvcache_host_async)
    check(fn(C.byref(p), _stream(k_cache.device)))
    return out_host


def host_pipeline_join(device: torch.device) -> None:
    """Make the current stream wait for every output copy of pipelined host calls still in flight."""
    check(lib.vattn_host_pipeline_join(_stream(device)))


def kernel_timing(op: int):
    """op 1 = start, 0 = stop/clear, 2 = read -> (total_ms, launches) of the dominant kernels."""
    ms, n = C.c_double(0.0), C.c_uint64(0)
    check(lib.vattn_kernel_timing(op, C.byref(ms), C.byref(n)))
    return ms.value, int(n.value)
