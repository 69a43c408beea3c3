"""Drop-in module shims: make the names the sarathi vAttention wrappers import resolve to this package.

The reference's wrapper files (sarathi/model_executor/attention/vattention_flashattention_wrapper.py,
vattention_flashinfer_wrapper.py, vattention_flashattention_pod_wrapper.py) and its cache engine
(sarathi/worker/cache_engine/vATTN_cache_engine.py) reach the hot path through five imports:

    import vattention                                         -> vattention_b200.vattention
    from flash_attn import flash_attn_with_kvcache, flash_attn_func
    from flashinfer import single_prefill_with_kv_cache       -> vattention_b200.attention
    import pod_attn as fused  (fused.true_fused_attn_with_kvcache)
    from sarathi.cache_ops import cache_flat

`install()` registers modules of those names in sys.modules whose attributes are this package's
operators, so that a sarathi process -- its wrapper and engine files UNMODIFIED -- runs on
libvattn_b200.so:

    import vattention_b200.dropin as dropin
    dropin.install()            # before sarathi's attention package is imported
    ... start the sarathi engine as usual ...

`installed()` is the same as a context manager that restores sys.modules afterwards (tests).
tests/test_dropin_reference_wrappers.py loads the reference's three wrapper files unmodified over
these shims.
"""
from __future__ import annotations

import contextlib
import sys
import types
from typing import Dict, Optional

SHIM_NAMES = ("vattention", "flash_attn", "flashinfer", "pod_attn", "sarathi.cache_ops")


def _flash_attn_func_unsupported(*_a, **_k):
    # imported by the wrappers (vattention_flashattention_wrapper.py:4) but never called on the
    # vAttention path
    raise RuntimeError("flash_attn_func is not on the vAttention hot path (the wrappers import it but call "
                       "flash_attn_with_kvcache only)")


def build_modules(ops=None, allocator=None) -> Dict[str, types.ModuleType]:
    """The shim modules, not yet registered.  `ops` / `allocator` default to vattention_b200.attention /
    vattention_b200.vattention; tests inject CPU stand-ins with the same call surface."""
    if ops is None:
        from . import attention as ops
    if allocator is None:
        from . import vattention as allocator
    mods: Dict[str, types.ModuleType] = {}

    va = types.ModuleType("vattention")
    va.__doc__ = "vattention_b200 shim of the reference's `vattention` torch extension (apis.h)"
    for name in ("init_kvcache", "reserve_physical_pages", "step", "step_async", "cleanup",
                 "alloc_new_batch_idx", "free_batch_idx", "num_free_kvblocks", "set_verbose",
                 "set_deferred_reclamation", "map_common_pages", "show_kvcache_config", "show_allocator_state"):
        if hasattr(allocator, name):
            setattr(va, name, getattr(allocator, name))
    mods["vattention"] = va

    fa = types.ModuleType("flash_attn")
    fa.flash_attn_with_kvcache = ops.flash_attn_with_kvcache
    fa.flash_attn_func = _flash_attn_func_unsupported
    fa.__version__ = "vattention_b200"
    mods["flash_attn"] = fa

    fi = types.ModuleType("flashinfer")
    fi.single_prefill_with_kv_cache = ops.single_prefill_with_kv_cache
    mods["flashinfer"] = fi

    pod = types.ModuleType("pod_attn")
    pod.true_fused_attn_with_kvcache = ops.true_fused_attn_with_kvcache
    pod.flash_attn_with_kvcache = ops.flash_attn_with_kvcache
    mods["pod_attn"] = pod

    co = types.ModuleType("sarathi.cache_ops")
    co.cache_flat = ops.cache_flat
    mods["sarathi.cache_ops"] = co
    return mods


def install(ops=None, allocator=None, override: bool = True) -> Dict[str, Optional[types.ModuleType]]:
    """Register the shims; returns what each name held before (None if absent).  With override=False
    names that are already importable are left alone."""
    previous: Dict[str, Optional[types.ModuleType]] = {}
    for name, mod in build_modules(ops, allocator).items():
        previous[name] = sys.modules.get(name)
        if override or name not in sys.modules:
            sys.modules[name] = mod
            if name == "sarathi.cache_ops" and "sarathi" in sys.modules:
                setattr(sys.modules["sarathi"], "cache_ops", mod)
    return previous


def uninstall(previous: Dict[str, Optional[types.ModuleType]]) -> None:
    for name, old in previous.items():
        if old is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = old


@contextlib.contextmanager
def installed(ops=None, allocator=None):
    previous = install(ops, allocator)
    try:
        yield
    finally:
        uninstall(previous)
