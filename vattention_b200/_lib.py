"""ctypes binding of libvattn_b200.so (the C ABI in include/vattn_b200.h).

The library is the product; there is no Python or CPU fallback behind it.  If it
is missing or a symbol is absent this module raises at import.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("VATTN_B200_LIB", _PKG / "libvattn_b200.so"))

OK = 0
ERR_INVALID, ERR_OOM, ERR_DRIVER, ERR_STATE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5
BACKEND_CUDA, BACKEND_HOST_MOCK = 0, 1
DTYPE_F16, DTYPE_BF16 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2


class VattnConfig(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "num_layers", "num_kv_heads", "head_size", "max_batch_size", "max_context_length",
        "bytes_per_elem", "page_size", "megacache", "tokens_per_page",
        "virt_buff_size_per_token", "virt_buff_size_per_req", "virt_buff_size",
        "max_pages_per_req", "phys_granularity", "num_tensors")]


class StepStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "critical_path_ns", "background_ns", "sync_pages_mapped", "async_pages_mapped",
        "driver_calls", "total_critical_path_ns", "total_background_ns", "max_background_ns",
        "total_sync_pages", "total_async_pages", "steps", "passes", "queued_steps")]


class FwdParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p),
        ("q_batch_stride", C.c_int64), ("q_row_stride", C.c_int64), ("q_head_stride", C.c_int64),
        ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("k_batch_stride", C.c_int64), ("k_row_stride", C.c_int64), ("k_head_stride", C.c_int64),
        ("v_batch_stride", C.c_int64), ("v_row_stride", C.c_int64), ("v_head_stride", C.c_int64),
        ("k_new", C.c_void_p), ("v_new", C.c_void_p),
        ("knew_batch_stride", C.c_int64), ("knew_row_stride", C.c_int64),
        ("knew_head_stride", C.c_int64),
        ("vnew_batch_stride", C.c_int64), ("vnew_row_stride", C.c_int64),
        ("vnew_head_stride", C.c_int64),
        ("out", C.c_void_p),
        ("o_batch_stride", C.c_int64), ("o_row_stride", C.c_int64), ("o_head_stride", C.c_int64),
        ("softmax_lse", C.c_void_p),
        ("cache_seqlens", C.c_void_p), ("cache_batch_idx", C.c_void_p),
        ("batch", C.c_int32), ("cache_batch", C.c_int32), ("seqlen_q", C.c_int32),
        ("seqlen_k", C.c_int32), ("seqlen_new", C.c_int32),
        ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("dtype", C.c_int32), ("causal", C.c_int32),
        ("softmax_scale", C.c_float),
        ("impl", C.c_int32), ("num_splits", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("rotary_cos", C.c_void_p), ("rotary_sin", C.c_void_p),
        ("rotary_dim", C.c_int32), ("rotary_interleaved", C.c_int32), ("seqlen_ro", C.c_int32),
    ]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m vattention_b200.build` "
            "(there is no fallback path)")
    return C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)


lib = _load()

_P = C.POINTER
_A = C.c_void_p  # opaque allocator handle

_SIGS = {
    "vattn_last_error": (C.c_char_p, []),
    "vattn_version": (C.c_char_p, []),
    "vattn_create": (C.c_int, [_P(_A), C.c_int]),
    "vattn_destroy": (C.c_int, [_A]),
    "vattn_init_kvcache": (C.c_int, [_A, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                     C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                     _P(C.c_uint64), _P(C.c_int), _P(C.c_int64), _P(C.c_int)]),
    "vattn_get_config": (C.c_int, [_A, _P(VattnConfig)]),
    "vattn_reserve_physical_pages": (C.c_int64, [_A, C.c_uint64]),
    "vattn_step": (C.c_int, [_A, _P(C.c_uint64), C.c_size_t, C.c_int]),
    "vattn_step_async": (C.c_int, [_A, _P(C.c_uint64), C.c_size_t]),
    "vattn_alloc_new_batch_idx": (C.c_int, [_A, C.c_uint64]),
    "vattn_free_batch_idx": (C.c_int, [_A, C.c_int]),
    "vattn_num_free_kvblocks": (C.c_uint64, [_A]),
    "vattn_cleanup": (C.c_int, [_A]),
    "vattn_set_verbose": (None, [_A, C.c_int]),
    "vattn_set_deferred_reclamation": (None, [_A, C.c_int]),
    "vattn_show_kvcache_config": (None, [_A]),
    "vattn_show_allocator_state": (None, [_A]),
    "vattn_map_common_pages": (C.c_int, [_A, C.c_uint64]),
    "vattn_wait_background": (C.c_int, [_A]),
    "vattn_set_compute_stream": (C.c_int, [_A, C.c_void_p, C.c_int]),
    "vattn_get_step_stats": (C.c_int, [_A, _P(StepStats)]),
    "vattn_get_state": (C.c_int, [_A, _P(C.c_uint64), _P(C.c_uint64), C.c_size_t]),
    "vattn_get_free_pool": (C.c_size_t, [_A, _P(C.c_uint64), C.c_size_t]),
    "vattn_get_pagemap": (C.c_size_t, [_A, _P(C.c_uint64), C.c_size_t]),
    "vattn_get_driver_log": (C.c_size_t, [_A, _P(C.c_uint64), C.c_size_t]),
    "vattn_clear_driver_log": (None, [_A]),
    "vattn_mock_set_capacity": (None, [_A, C.c_uint64]),
    "vattn_mock_set_call_delay_us": (None, [_A, C.c_uint64]),
    "vattn_mock_fence_counts": (None, [_A, C.POINTER(C.c_uint64)]),
    "vattn_set_queueing": (C.c_int, [_A, C.c_int]),
    "vattn_fwd_kvcache_workspace": (C.c_size_t, [_P(FwdParams)]),
    "vattn_fwd_kvcache": (C.c_int, [_P(FwdParams), C.c_void_p]),
    "vattn_single_prefill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_size_t,
                                       C.c_void_p]),
    "vattn_pod_workspace": (C.c_size_t, [_P(FwdParams), _P(FwdParams)]),
    "vattn_pod_fwd": (C.c_int, [_P(FwdParams), _P(FwdParams), C.c_int32, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    "vattn_cache_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                   C.c_int32, C.c_void_p]),
    "vattn_fwd_kvcache_host": (C.c_int, [_P(FwdParams), C.c_void_p]),
    "vattn_fwd_kvcache_host_async": (C.c_int, [_P(FwdParams), C.c_void_p]),
    "vattn_fwd_kvcache_host_pipelined": (C.c_int, [_P(FwdParams), C.c_void_p]),
    "vattn_host_pipeline_join": (C.c_int, [C.c_void_p]),
    "vattn_allreduce_oneshot": (C.c_int, [_P(C.c_uint64), _P(C.c_uint64), C.c_void_p, C.c_int64, C.c_int,
                                          C.c_int, C.c_int, C.c_uint32, C.c_void_p]),
    "vattn_oproj_allreduce_recv_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "vattn_oproj_allreduce_flag_bytes": (C.c_size_t, [C.c_int32]),
    "vattn_oproj_allreduce": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, _P(C.c_uint64), _P(C.c_uint64),
                                        C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vattn_launch_count": (C.c_uint64, []),
    "vattn_kernel_timing": (C.c_int, [C.c_int, _P(C.c_double), _P(C.c_uint64)]),
    "vattn_selftest_umma": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == library/header mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.vattn_last_error() or b"").decode()


def check(status: int) -> int:
    """Raise the library's error as RuntimeError (what the reference's pybind layer does
    for std::runtime_error, e.g. vattention.cu:295)."""
    if status < 0:
        raise RuntimeError(last_error() or f"vattn error {status}")
    return status
