"""Drop-in mirror of the reference's `vattention` extension module.

Same 13 module-level functions, argument order and error behaviour as the pybind
module registered at /root/reference/vattention/vattention.cu:614-637 (bodies in
apis.h:1-63); the only in-tree caller, and therefore the contract, is
sarathi-lean/sarathi/worker/cache_engine/vATTN_cache_engine.py:25-195.

    import vattention_b200.vattention as vattention      # instead of `import vattention`

All state lives in libvattn_b200.so behind the C ABI (include/vattn_b200.h); this
file only marshals arguments and wraps the returned virtual addresses in torch
tensors (the reference does that with a custom at::Allocator, vtensor.h:9-125).
One process-global allocator instance, like the reference's
`static vAttentionCachingAllocator vattn;` (apis.h:1).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from ._lib import lib, check

_handle: Optional[C.c_void_p] = None
_backend = _lib.BACKEND_CUDA
_tensors: List[torch.Tensor] = []   # kept alive like k_tensors/v_tensors (utils.h:37-38)


def _use_backend(backend: int) -> None:
    """Test hook: select the HOST_MOCK driver before init_kvcache (CPU bookkeeping tests)."""
    global _backend, _handle
    if _handle is not None:
        check(lib.vattn_destroy(_handle))
        _handle = None
    _backend = backend


def _get() -> C.c_void_p:
    global _handle
    if _handle is None:
        h = C.c_void_p()
        check(lib.vattn_create(C.byref(h), _backend))
        _handle = h
    return _handle


class _VirtualBuffer:
    """Exposes a raw device VA range through __cuda_array_interface__ so torch can adopt it
    without an extension module; the storage owns nothing (vtensor.h:45,59-62)."""

    def __init__(self, ptr: int, nelem: int, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": (nelem,), "typestr": typestr, "data": (ptr, False), "version": 3,
            "strides": None,
        }


_CARRIER = {2: ("<i2", torch.int16), 4: ("<i4", torch.int32), 1: ("|u1", torch.uint8),
            8: ("<i8", torch.int64)}


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int),
                ("dtype", _DLDataType), ("shape", C.POINTER(C.c_int64)),
                ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p),
                             ("deleter", C.c_void_p)]
_DL_CODES = {torch.float16: (2, 16), torch.bfloat16: (4, 16), torch.float32: (2, 32),
             torch.int16: (0, 16), torch.uint8: (1, 8), torch.int8: (0, 8)}
_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]


def _tensor_from_va_dlpack(ptr: int, shape, dtype: torch.dtype, device: int,
                           device_type: int = 2) -> torch.Tensor:
    """Build a DLManagedTensor by hand.  The descriptor is malloc'ed and never freed (a few
    dozen bytes per cache tensor) so it outlives every tensor that aliases it, including
    at interpreter teardown; the deleter is NULL because the allocator, not the tensor,
    owns the address range (the reference's DataPtr deleter is a no-op too, vtensor.h:59-62)."""
    code, bits = _DL_CODES[dtype]
    nd = len(shape)
    raw = _libc.malloc(C.sizeof(_DLManagedTensor) + 8 * nd)
    mt = _DLManagedTensor.from_address(raw)
    shp = (C.c_int64 * nd).from_address(raw + C.sizeof(_DLManagedTensor))
    for i, s in enumerate(shape):
        shp[i] = int(s)
    mt.dl_tensor.data = ptr
    mt.dl_tensor.device = _DLDevice(device_type, device)  # 2 = kDLCUDA, 1 = kDLCPU
    mt.dl_tensor.ndim = nd
    mt.dl_tensor.dtype = _DLDataType(code, bits, 1)
    mt.dl_tensor.shape = C.cast(raw + C.sizeof(_DLManagedTensor), C.POINTER(C.c_int64))
    mt.dl_tensor.strides = None                 # compact row-major
    mt.dl_tensor.byte_offset = 0
    mt.manager_ctx = None
    mt.deleter = None
    new_capsule = C.pythonapi.PyCapsule_New
    new_capsule.restype = C.py_object
    new_capsule.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    cap = new_capsule(raw, b"dltensor", None)
    return torch.utils.dlpack.from_dlpack(cap)


def _tensor_from_va(ptr: int, shape, dtype: torch.dtype, device: int) -> torch.Tensor:
    """Adopt a reserved (possibly still unmapped) device VA range as a torch tensor.  DLPack
    carries the device explicitly, so torch never has to inspect the pointer; the
    __cuda_array_interface__ route is the fallback."""
    try:
        t = _tensor_from_va_dlpack(ptr, shape, dtype, device)
    except Exception:  # pragma: no cover - depends on the torch build
        itemsize = torch.empty((), dtype=dtype).element_size()
        n = 1
        for s in shape:
            n *= int(s)
        typestr, _carrier = _CARRIER[itemsize]
        buf = _VirtualBuffer(ptr, n, typestr)
        t = torch.as_tensor(buf, device=torch.device("cuda", device))
        t = t.view(dtype).view(*[int(s) for s in shape])
    if t.data_ptr() != ptr or t.dtype != dtype:
        raise RuntimeError("[vattn] failed to wrap the virtual tensor")
    return t


def init_kvcache(num_layers: int, num_kv_heads: int, head_size: int, max_batch_size: int,
                 max_context_length: int, device: int, dtype, page_size: int,
                 megacache: bool) -> List[torch.Tensor]:
    """apis.h:3-13.  Returns [K_0..K_{L-1}, V_0..V_{L-1}] of shape [B, maxlen, Hkv, D], or
    [K, V] of shape [B, maxlen, L, Hkv, D] with megacache (vattention.cu:142-187).  The
    tensors are backed by reserved virtual address space only."""
    h = _get()
    itemsize = int(getattr(dtype, "itemsize", 0)) or torch.empty((), dtype=dtype).element_size()
    n_max = 2 * int(num_layers)
    ptrs = (C.c_uint64 * max(n_max, 2))()
    n = C.c_int(0)
    shape = (C.c_int64 * 5)()
    ndim = C.c_int(0)
    check(lib.vattn_init_kvcache(h, num_layers, num_kv_heads, head_size, max_batch_size,
                                 max_context_length, int(device), itemsize, int(page_size),
                                 1 if megacache else 0, ptrs, C.byref(n), shape, C.byref(ndim)))
    shp = [shape[i] for i in range(ndim.value)]
    if _backend == _lib.BACKEND_HOST_MOCK:
        # no device memory behind a mock VA: hand back meta tensors of the right geometry
        out = [torch.empty(shp, dtype=dtype, device="meta") for _ in range(n.value)]
    else:
        out = [_tensor_from_va(int(ptrs[i]), shp, dtype, int(device)) for i in range(n.value)]
    _tensors[:] = out
    return list(out)


def reserve_physical_pages(free_memory: int) -> int:
    """apis.h:23-25"""
    r = lib.vattn_reserve_physical_pages(_get(), int(free_memory))
    if r < 0:
        check(int(r))
    return int(r)


def _lens(seq_lens) -> "C.Array":
    return (C.c_uint64 * len(seq_lens))(*[int(x) for x in seq_lens])


def step(seq_lens, eager_reclaim: bool) -> None:
    """apis.h:27-29 (all mapping synchronous; the `_sync` backends)"""
    arr = _lens(seq_lens)
    check(lib.vattn_step(_get(), arr, len(seq_lens), 1 if eager_reclaim else 0))


def step_async(seq_lens) -> None:
    """apis.h:31-35.  ctypes drops the GIL for the duration of the call, as the reference
    does with Py_BEGIN_ALLOW_THREADS."""
    arr = _lens(seq_lens)
    check(lib.vattn_step_async(_get(), arr, len(seq_lens)))


def alloc_new_batch_idx(seqlen: int) -> int:
    """apis.h:53-55; -1 when no request slot is free (caller asserts)."""
    r = lib.vattn_alloc_new_batch_idx(_get(), int(seqlen))
    if r < -1:
        check(r + 100)
    return r


def free_batch_idx(reqId: int) -> None:
    """apis.h:57-59"""
    check(lib.vattn_free_batch_idx(_get(), int(reqId)))


def num_free_kvblocks() -> int:
    """apis.h:61-63"""
    return int(lib.vattn_num_free_kvblocks(_get()))


def cleanup() -> None:
    """apis.h:41-43"""
    global _handle
    if _handle is not None:
        check(lib.vattn_cleanup(_handle))
    _tensors.clear()


def set_verbose(val: bool) -> None:
    """apis.h:37-39"""
    lib.vattn_set_verbose(_get(), 1 if val else 0)


def set_deferred_reclamation(val: bool) -> None:
    """apis.h:45-47"""
    lib.vattn_set_deferred_reclamation(_get(), 1 if val else 0)


def show_kvcache_config() -> None:
    """apis.h:15-17"""
    lib.vattn_show_kvcache_config(_get())


def show_allocator_state() -> None:
    """apis.h:19-21"""
    lib.vattn_show_allocator_state(_get())


def map_common_pages(num_tokens: int) -> None:
    """apis.h:49-51"""
    check(lib.vattn_map_common_pages(_get(), int(num_tokens)))


# ---- additions beyond the reference surface (used by tests / bench) -------------------

def wait_background() -> None:
    check(lib.vattn_wait_background(_get()))


def set_compute_stream(stream: Optional[int], enable: bool = True) -> None:
    check(lib.vattn_set_compute_stream(_get(), C.c_void_p(stream or 0), 1 if enable else 0))


def get_config() -> dict:
    cfg = _lib.VattnConfig()
    check(lib.vattn_get_config(_get(), C.byref(cfg)))
    return {n: int(getattr(cfg, n)) for n, _ in cfg._fields_}


def get_step_stats() -> dict:
    st = _lib.StepStats()
    check(lib.vattn_get_step_stats(_get(), C.byref(st)))
    return {n: int(getattr(st, n)) for n, _ in st._fields_}


def get_state() -> dict:
    """mapped_pages / seq_lens / free pool / page map, for bit-exact parity tests."""
    h = _get()
    b = get_config()["max_batch_size"]
    mapped = (C.c_uint64 * b)()
    lens = (C.c_uint64 * b)()
    check(lib.vattn_get_state(h, mapped, lens, b))
    n = lib.vattn_get_free_pool(h, None, 0)
    pool = (C.c_uint64 * max(n, 1))()
    lib.vattn_get_free_pool(h, pool, n)
    m = lib.vattn_get_pagemap(h, None, 0)
    words = (C.c_uint64 * max(5 * m, 1))()
    lib.vattn_get_pagemap(h, words, m)
    return {
        "mapped_pages": [int(x) for x in mapped],
        "seq_lens": [int(x) for x in lens],
        "pool": [int(pool[i]) for i in range(n)],
        "pagemap": [[int(words[5 * i + j]) for j in range(5)] for i in range(m)],
        "num_free_kvblocks": num_free_kvblocks(),
    }


def get_driver_log() -> list:
    h = _get()
    n = lib.vattn_get_driver_log(h, None, 0)
    words = (C.c_uint64 * max(4 * n, 1))()
    lib.vattn_get_driver_log(h, words, n)
    return [tuple(int(words[4 * i + j]) for j in range(4)) for i in range(n)]


def clear_driver_log() -> None:
    lib.vattn_clear_driver_log(_get())


def set_queueing(on: bool) -> None:
    """step_async may ride behind the mapper pass in flight (default) or always wait for it (the reference)."""
    check(lib.vattn_set_queueing(_get(), int(bool(on))))


def mock_set_call_delay_us(us: int) -> None:
    """HOST_MOCK: every map / set_access / unmap takes `us` microseconds."""
    lib.vattn_mock_set_call_delay_us(_get(), int(us))


def mock_fence_counts() -> dict:
    out = (C.c_uint64 * 4)()
    lib.vattn_mock_fence_counts(_get(), out)
    return {"records": [int(out[0]), int(out[1])], "waits": [int(out[2]), int(out[3])]}


def mock_set_capacity(nbytes: int) -> None:
    """HOST_MOCK only: physical bytes the mock device can hold (0 = unlimited)."""
    lib.vattn_mock_set_capacity(_get(), int(nbytes))
