"""Deterministic allocator API traces + replay helpers (TEST INFRASTRUCTURE ONLY).

A trace is a JSON-able list of ops (format in oracle/ref_driver.py).  The same trace is
replayed on (a) the reference extension on the GPU box (ref_driver.py), (b) the oracle
(allocator_model.py) and (c) the product allocator (vattention_b200.vattention); snapshots
after every op must agree.  The call pattern follows the allocator's only in-tree caller,
vATTNCacheEngine.step / on_step_completion (vATTN_cache_engine.py:91-152): new sequences
take a reqId with alloc_new_batch_idx(len), every iteration passes the full curr_seq_lens
vector to step/step_async, finished sequences free their reqId.
"""
from __future__ import annotations

import random
from typing import List

from .allocator_model import MB, AllocatorModel, AllocatorOOM


def build_trace(name: str) -> List[list]:
    cfgs = {
        # name: (L, Hkv, D, B, ctx, dtype, page, mega, pool_pages, mode, seed, steps)
        "llama8b_async": (2, 8, 128, 8, 32768, "bf16", 2 * MB, 0, 176, "async", 0, 120),
        "yi6b_sync": (3, 4, 128, 6, 65536, "fp16", 2 * MB, 0, 96, "sync", 1, 100),
        "tp8_async_nodefer": (2, 1, 128, 6, 131072, "bf16", 2 * MB, 0, 80, "async_nodefer", 2, 100),
        "mega_async": (4, 2, 128, 6, 32768, "bf16", 2 * MB, 1, 32, "async", 3, 120),
        "tight_pool_sync": (2, 8, 128, 8, 16384, "bf16", 2 * MB, 0, 40, "sync", 4, 150),
    }
    L, Hkv, D, B, ctx, dt, page, mega, pool, mode, seed, steps = cfgs[name]
    rng = random.Random(seed)
    itemsize = 2
    tpp = page // (Hkv * D * itemsize * (L if mega else 1))
    trace = [["init", L, Hkv, D, B, ctx, dt, page, mega], ["reserve", pool * page]]
    if mode == "async_nodefer":
        trace.append(["set_deferred", False])
    lens = [0] * B
    # the trace generator tracks only what the caller knows (its own curr_seq_lens); reqIds
    # come back from the allocator at replay time, so "alloc" is followed by a marker that
    # tells the replayer to write the new length into the slot it was given
    for _ in range(steps):
        r = rng.random()
        active = [i for i in range(B) if lens[i]]
        if r < 0.22 and len(active) < B:
            n = rng.choice([1, tpp - 1, tpp, tpp + 1, 2 * tpp + 9, rng.randrange(1, ctx // 3)])
            trace.append(["alloc", n])
            lens[_first_free(lens)] = n  # placeholder slot; replay uses the real reqId
        elif r < 0.32 and active:
            trace.append(["free_active", rng.randrange(len(active))])
            lens[active[rng.randrange(len(active))]] = 0
        elif r < 0.38:
            trace.append(["nfree"])
        else:
            trace.append(["decode", rng.randrange(1, 4), rng.random() < 0.2, mode])
    return trace


def _first_free(lens):
    for i, v in enumerate(lens):
        if v == 0:
            return i
    return 0


TRACE_NAMES = ["llama8b_async", "yi6b_sync", "tp8_async_nodefer", "mega_async", "tight_pool_sync"]


def expand(trace: List[list], backend) -> List[list]:
    """Resolve the caller-side ops (`alloc`, `free_active`, `decode`) into concrete API calls by
    running them against `backend` (anything with the allocator API).  Returns the concrete
    trace (only ops ref_driver.py understands) -- identical for every conforming backend,
    which is itself part of the check."""
    concrete = []
    cfg = trace[0]
    B, ctx = cfg[4], cfg[5]
    page, mega, L, Hkv, D = cfg[7], cfg[8], cfg[1], cfg[2], cfg[3]
    tpp = page // (Hkv * D * 2 * (L if mega else 1))
    lens = [0] * B
    rng = random.Random(1234)
    for op in trace:
        if op[0] in ("init", "reserve", "set_deferred", "nfree"):
            backend(op)
            concrete.append(op)
        elif op[0] == "alloc":
            rid = backend(["alloc", op[1]])
            concrete.append(["alloc", op[1]])
            if rid is not None and rid >= 0:
                lens[rid] = op[1]
        elif op[0] == "free_active":
            active = [i for i in range(B) if lens[i]]
            if not active:
                continue
            rid = active[op[1] % len(active)]
            backend(["free", rid])
            concrete.append(["free", rid])
            lens[rid] = 0
        elif op[0] == "decode":
            _, n, jump, mode = op
            for _ in range(n):
                for i in range(B):
                    if lens[i] and lens[i] < ctx - 1:
                        lens[i] = min(ctx - 1, lens[i] + (tpp // 2 if jump and rng.random() < 0.3 else 1))
                call = ["step", list(lens), True] if mode == "sync" else ["step_async", list(lens)]
                err = backend(call)
                concrete.append(call)
                if err == "oom":
                    big = max(range(B), key=lambda i: lens[i])
                    lens[big] = 0
                    backend(["free", big])
                    concrete.append(["free", big])
    return concrete


class ModelBackend:
    """Replays concrete ops on the oracle; collects snapshots in ref_driver.py's format."""

    def __init__(self):
        self.m = None
        self.snaps = []

    def __call__(self, op):
        ret, err = None, None
        try:
            if op[0] == "init":
                _, L, Hkv, D, B, ctx, dt, page, mega = op
                self.m = AllocatorModel(L, Hkv, D, B, ctx, 2, page, bool(mega))
                shape = [B, ctx, L, Hkv, D] if mega else [B, ctx, Hkv, D]
                stride = [1] * len(shape)
                for i in range(len(shape) - 2, -1, -1):
                    stride[i] = stride[i + 1] * shape[i + 1]
                ret = [2 if mega else 2 * L, shape, stride]
            elif op[0] == "reserve":
                ret = self.m.reserve_physical_pages(op[1])
            elif op[0] == "step":
                self.m.step(op[1], bool(op[2]))
            elif op[0] == "step_async":
                self.m.step_async(op[1])
            elif op[0] == "alloc":
                ret = self.m.alloc_new_batch_idx(op[1])
            elif op[0] == "free":
                self.m.free_batch_idx(op[1])
            elif op[0] == "nfree":
                ret = self.m.num_free_kvblocks()
            elif op[0] == "set_deferred":
                self.m.set_deferred_reclamation(bool(op[1]))
        except AllocatorOOM as e:
            err = str(e)
        s = self.m.snapshot()
        self.snaps.append({"op": op, "ret": ret, "err": err, "mapped_pages": s["mapped_pages"],
                           "seq_lens": s["seq_lens"], "pool": s["pool"],
                           "num_free_kvblocks": s["num_free_kvblocks"]})
        if err:
            return "oom"
        return ret


class ProductBackend:
    """Replays concrete ops on vattention_b200.vattention (mock or CUDA backend)."""

    def __init__(self, va, torch):
        self.va, self.torch, self.snaps = va, torch, []

    def __call__(self, op):
        va, ret, err = self.va, None, None
        try:
            if op[0] == "init":
                _, L, Hkv, D, B, ctx, dt, page, mega = op
                dtype = {"bf16": self.torch.bfloat16, "fp16": self.torch.float16}[dt]
                # the tunable outlives cleanup() (it is a process global in the reference too,
                # utils.h:78); every trace starts from the documented default
                va.set_deferred_reclamation(True)
                ts = va.init_kvcache(L, Hkv, D, B, ctx, 0, dtype, page, bool(mega))
                ret = [len(ts), list(ts[0].shape), list(ts[0].stride())]
            elif op[0] == "reserve":
                ret = va.reserve_physical_pages(op[1])
            elif op[0] == "step":
                va.step(op[1], bool(op[2]))
            elif op[0] == "step_async":
                va.step_async(op[1])
                va.wait_background()
            elif op[0] == "alloc":
                ret = va.alloc_new_batch_idx(op[1])
            elif op[0] == "free":
                va.free_batch_idx(op[1])
            elif op[0] == "nfree":
                ret = va.num_free_kvblocks()
            elif op[0] == "set_deferred":
                va.set_deferred_reclamation(bool(op[1]))
        except RuntimeError as e:
            err = str(e).splitlines()[0]
            va.set_verbose(False)
        s = va.get_state()
        self.snaps.append({"op": op, "ret": ret, "err": err, "mapped_pages": s["mapped_pages"],
                           "seq_lens": s["seq_lens"], "pool": s["pool"],
                           "num_free_kvblocks": s["num_free_kvblocks"]})
        if err:
            return "oom"
        return ret


def compare_snaps(got: list, want: list, what: str) -> None:
    assert len(got) == len(want), f"{what}: {len(got)} vs {len(want)} snapshots"
    for i, (g, w) in enumerate(zip(got, want)):
        for key in ("ret", "mapped_pages", "seq_lens", "pool", "num_free_kvblocks"):
            assert g[key] == w[key], f"{what}: op #{i} {w['op'][:2]} field {key}: {g[key]} != {w[key]}"
        assert (g["err"] is None) == (w["err"] is None), f"{what}: op #{i} error mismatch: {g['err']} vs {w['err']}"
        if w["err"]:
            assert "OOM on demand" in g["err"] and "OOM on demand" in w["err"]
