#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs[2], configs[3]) and the allocator-overlap
measurement.  bench.py carries the headline decode metric; this script produces the numbers
DESIGN.md / profiles/ quote for the other rows of SURVEY 8(d).  One JSON line per workload.

  python scripts/bench_extra.py prefill [--chunk 2048] [--impl ours|fa|fi]   configs[2]
  python scripts/bench_extra.py pod [--impl ours|fa]                         configs[3]
  python scripts/bench_extra.py alloc                                        step_async overlap
  python scripts/bench_extra.py decode [--ragged] [--dtype fp16] [--hq 8 --hkv 1 --batch 16 --ctx 131072]
                                                                             configs[1] variants, config 5 per-GPU shapes
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from vattention_b200 import attention as att  # noqa: E402
from vattention_b200 import vattention as va  # noqa: E402

DEV = torch.device("cuda", 0)
PAGE = 2 << 20


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    d = json.loads(f.read_text()) if f.exists() else {}
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0)


def timed(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def prefill(args):
    """configs[2]: Yi-6B-200K shapes (Hq 32, Hkv 4, D 128), chunked prefill of a 128K context, one
    layer, K/V in a vAttention tensor.  Per chunk: cache_flat(k, v) + causal attention of the chunk
    over everything cached so far.  FLOPs (causal-exact) = 4*Hq*D*(c*p + c(c+1)/2) per chunk."""
    Hq, Hkv, D, S, c = 32, 4, 128, args.ctx, args.chunk
    dtype = torch.bfloat16
    torch.zeros(1, device=DEV)
    kc, vc = va.init_kvcache(1, Hkv, D, 1, S, 0, dtype, args.page_kb << 10, False)
    va.reserve_physical_pages(2 * S * Hkv * D * 2 + (8 << 20))
    va.step([S], True)
    g = torch.Generator(device=DEV).manual_seed(0)
    q = torch.randn(S, Hq, D, device=DEV, generator=g).to(dtype)
    k = torch.randn(S, Hkv, D, device=DEV, generator=g).to(dtype)
    v = torch.randn(S, Hkv, D, device=DEV, generator=g).to(dtype)
    if args.impl == "fa":
        from flash_attn import flash_attn_with_kvcache as fa_fwd
    if args.impl == "fi":
        from flashinfer import single_prefill_with_kv_cache as fi_fwd

    def whole_prefill():
        for p in range(0, S, c):
            att.cache_flat(k[p:p + c], v[p:p + c], kc[0][p:], vc[0][p:], "auto")
            total = torch.tensor([p + c], dtype=torch.int32, device=DEV)
            if args.impl == "ours":
                att.flash_attn_with_kvcache(q[p:p + c].unsqueeze(0), kc, vc, cache_seqlens=total, causal=True)
            elif args.impl == "fa":
                fa_fwd(q[p:p + c].unsqueeze(0), kc, vc, cache_seqlens=total, causal=True)
            else:
                fi_fwd(q[p:p + c], kc[0][:p + c], vc[0][:p + c], causal=True)

    att.kernel_timing(1)
    ms = timed(whole_prefill, 1, args.iters)
    kern_ms, kern_n = att.kernel_timing(2)
    att.kernel_timing(0)
    flops = sum(4 * Hq * D * (c * p + c * (c + 1) // 2) for p in range(0, S, c))
    _, burst, sustained = peaks()
    tf = flops / (ms * 1e-3) / 1e12
    out = {"workload": f"prefill ctx{S} chunk{c} Yi-6B shapes bf16, page {args.page_kb} KB", "impl": args.impl,
           "ms_per_prefill": round(ms, 2), "tflops": round(tf, 1), "flops": flops,
           "frac_of_measured_bf16_burst": round(tf / burst, 4),
           "frac_of_measured_bf16_sustained": round(tf / sustained, 4)}
    if kern_n and args.impl == "ours":
        per_pass = kern_ms / (args.iters + 1)   # timing covered the warm-up pass too
        out["attention_kernel_ms_per_prefill"] = round(per_pass, 2)
        out["attention_kernel_tflops"] = round(flops / (per_pass * 1e-3) / 1e12, 1)
    va.cleanup()
    print(json.dumps(out))


def pod(args):
    """configs[3]: Llama-3-8B shapes fp16, 8 prefills (16K queries over 16K keys) + 56 decodes at 4K
    context, one layer.  fused call vs the two calls back to back."""
    Hq, Hkv, D = 32, 8, 128
    Bp, Sp, Bd, Sd = args.prefills, args.prefill_len, args.decodes, args.decode_len
    dtype = torch.float16
    g = torch.Generator(device=DEV).manual_seed(0)
    Sq = args.prefill_chunk or Sp       # wrapper-realistic variant: one chunk deep in a long context
    q_p = torch.randn(Bp, Sq, Hq, D, device=DEV, generator=g).to(dtype)
    kc_p = torch.randn(Bp, Sp, Hkv, D, device=DEV, generator=g).to(dtype)
    vc_p = torch.randn(Bp, Sp, Hkv, D, device=DEV, generator=g).to(dtype)
    lens_p = torch.full((Bp,), Sp, dtype=torch.int32, device=DEV)
    q_d = torch.randn(Bd, 1, Hq, D, device=DEV, generator=g).to(dtype)
    kc_d = torch.randn(Bd, Sd, Hkv, D, device=DEV, generator=g).to(dtype)
    vc_d = torch.randn(Bd, Sd, Hkv, D, device=DEV, generator=g).to(dtype)
    kn = torch.randn(Bd, 1, Hkv, D, device=DEV, generator=g).to(dtype)
    vn = torch.randn(Bd, 1, Hkv, D, device=DEV, generator=g).to(dtype)
    lens_d = torch.full((Bd,), Sd - 1, dtype=torch.int32, device=DEV)
    idx = torch.randperm(Bd, device=DEV, generator=g).int()
    if args.impl == "fa":
        from flash_attn import flash_attn_with_kvcache as fwd
    else:
        fwd = att.flash_attn_with_kvcache

    def run_p():
        return fwd(q_p, kc_p, vc_p, cache_seqlens=lens_p, causal=True)

    def run_d():
        return fwd(q_d, kc_d, vc_d, kn, vn, cache_seqlens=lens_d, cache_batch_idx=idx, causal=True)

    def run_serial():
        run_p()
        run_d()

    def run_fused(fp=15):
        att.true_fused_attn_with_kvcache(q_p, kc_p, vc_p, q_d, kc_d, vc_d, kn, vn, causal=True,
                                         cache_seqlens_p=lens_p, cache_seqlens_d=lens_d,
                                         cache_batch_idx=idx, fused_params=fp)

    side = torch.cuda.Stream(device=DEV)

    def run_two_streams():              # attn_sweep.py:63-65's third arm
        side.wait_stream(torch.cuda.current_stream(DEV))
        with torch.cuda.stream(side):
            run_d()
        run_p()
        torch.cuda.current_stream(DEV).wait_stream(side)

    t_p, t_d = timed(run_p, 1, args.iters), timed(run_d, 3, 20)
    t_s = timed(run_serial, 1, args.iters)
    t_2 = timed(run_two_streams, 1, args.iters)
    out = {"workload": f"pod {Bp}x prefill {Sq}q@{Sp} + {Bd}x decode@{Sd} Llama-3-8B fp16", "impl": args.impl,
           "prefill_ms": round(t_p, 3), "decode_ms": round(t_d, 4), "serial_ms": round(t_s, 3),
           "two_streams_ms": round(t_2, 3)}
    flops = Bp * 4 * Hq * D * (Sq * (Sp - Sq) + Sq * (Sq + 1) // 2)
    dbytes = 2 * 2 * Hkv * D * Bd * Sd
    hbm, burst, _ = peaks()
    out["prefill_tflops"] = round(flops / (t_p * 1e-3) / 1e12, 1)
    out["decode_gbps"] = round(dbytes / (t_d * 1e-3) / 1e9, 1)
    out["roofline_ms"] = round(max(flops / (burst * 1e12), dbytes / (hbm * 1e9)) * 1e3, 3)
    if args.impl == "ours":
        # true_fused_attn_with_kvcache: 15 = auto (co-scheduled specialised kernels), 9 = the
        # persistent single kernel
        t_f = timed(run_fused, 1, args.iters)
        t_k = timed(lambda: run_fused(9), 1, args.iters)
        out["pod_call_ms"] = round(t_f, 3)
        out["pod_call_vs_serial"] = round(t_s / t_f, 4)
        out["persistent_kernel_ms"] = round(t_k, 3)
        out["persistent_kernel_vs_serial"] = round(t_s / t_k, 4)
        # the dual-role kernel (fused_params 64): a prefill and a decode pipeline in every CTA
        t_l = timed(lambda: run_fused(64), 1, args.iters)
        out["dual_role_kernel_ms"] = round(t_l, 3)
        out["dual_role_kernel_vs_serial"] = round(t_s / t_l, 4)
    print(json.dumps(out))


def alloc(args):
    """How much page-mapping latency lands on the critical path.  A decode batch whose sequences
    cross a 2 MB page boundary at different steps (ragged lengths): per step, time the step_async
    call (critical path) and the mapper thread's pass, with the decode kernels of 32 layer-calls in
    flight; compare with synchronous `step` (all mapping on the critical path)."""
    L, Hkv, Hq, D, B, ctx = 32, 8, 32, 128, 16, 32768
    dtype = torch.bfloat16
    torch.zeros(1, device=DEV)
    res = {}
    for mode in ("async", "sync"):
        ts = va.init_kvcache(L, Hkv, D, B, ctx, 0, dtype, PAGE, False)
        va.reserve_physical_pages(B * 6 * 2 * L * PAGE)
        tpp = va.get_config()["tokens_per_page"]
        # sequence b sits (b+1)*3 tokens before a page boundary: one crossing every 3 steps
        lens = [2 * tpp - 3 * (b + 1) for b in range(B)]
        (va.step_async if mode == "async" else lambda l: va.step(l, True))(lens)
        va.wait_background()
        q = torch.randn(B, 1, Hq, D, device=DEV).to(dtype)
        kn = torch.randn(B, 1, Hkv, D, device=DEV).to(dtype)
        idx = torch.arange(B, device=DEV).int()
        crit, bg, sync_pages, async_pages, wall = [], [], 0, 0, []
        for _ in range(args.alloc_steps):
            lens = [n + 1 for n in lens]
            t0 = time.perf_counter()
            if mode == "async":
                va.step_async(lens)
            else:
                va.step(lens, True)
            t1 = time.perf_counter()
            sl = torch.tensor([n - 1 for n in lens], dtype=torch.int32, device=DEV)
            for layer in range(L):
                att.flash_attn_with_kvcache(q, ts[layer][:, :max(lens)], ts[L + layer][:, :max(lens)], kn, kn,
                                            cache_seqlens=sl, cache_batch_idx=idx, causal=True)
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
            crit.append(t1 - t0)
            if mode == "async":
                va.wait_background()
            st = va.get_step_stats()
            bg.append(st["background_ns"] * 1e-9)
            sync_pages += st["sync_pages_mapped"]
            async_pages += st["async_pages_mapped"] if mode == "async" else 0
        res[mode] = {"critical_path_us_mean": round(1e6 * sum(crit) / len(crit), 1),
                     "critical_path_us_max": round(1e6 * max(crit), 1),
                     "background_pass_us_mean": round(1e6 * sum(bg) / len(bg), 1),
                     "background_pass_us_max": round(1e6 * max(bg), 1),
                     "iteration_ms_mean": round(1e3 * sum(wall) / len(wall), 3),
                     "pages_mapped_on_critical_path": sync_pages, "pages_mapped_in_background": async_pages}
        va.cleanup()
    print(json.dumps({"workload": f"alloc overlap: B{B} L{L} Llama-3-8B, a page-boundary crossing every 3 steps, "
                                  f"{args.alloc_steps} steps", **res}))


def decode(args):
    """configs[1] variants and config 5's per-GPU shapes: ONE layer-call of decode attention with the
    fused k/v append over vAttention tensors (num_layers = 1, as bench_pagesize.py does).  Bytes =
    SURVEY 8(d): K and V once (+1 appended row), Q read, O write, k/v new write."""
    Hq, Hkv, D, B, ctx = args.hq, args.hkv, 128, args.batch, args.ctx
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    torch.zeros(1, device=DEV)
    kc, vc = va.init_kvcache(1, Hkv, D, B, ctx, 0, dtype, PAGE, False)
    va.reserve_physical_pages(2 * B * ctx * Hkv * D * 2 + (8 << 20))
    g = torch.Generator().manual_seed(0)
    if args.ragged:
        lens = torch.randint(ctx // 2, ctx, (B,), generator=g).tolist()
    else:
        lens = [ctx - 1] * B
    va.step([n + 1 for n in lens], True)
    gd = torch.Generator(device=DEV).manual_seed(0)
    for b in range(B):
        kc[b, :lens[b]].normal_(generator=gd)
        vc[b, :lens[b]].normal_(generator=gd)
    q = torch.randn(B, 1, Hq, D, device=DEV, generator=gd).to(dtype)
    kn = torch.randn(B, 1, Hkv, D, device=DEV, generator=gd).to(dtype)
    vn = torch.randn(B, 1, Hkv, D, device=DEV, generator=gd).to(dtype)
    sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    idx = torch.arange(B, device=DEV).int()
    mx = max(lens) + 1
    if args.impl == "fa":
        from flash_attn import flash_attn_with_kvcache as fwd
    else:
        fwd = att.flash_attn_with_kvcache
    # every layer-call reads 2*B*ctx*Hkv*D*2 bytes; below the 126 MB L2 rotate over nothing would be
    # wrong, so small shapes are flushed by a 256 MB write between calls (outside the events)
    kv_bytes = 2 * 2 * Hkv * D * sum(n + 1 for n in lens)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV) if kv_bytes < (512 << 20) else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.calls)]
    for i in range(3 + args.calls):
        if flush is not None:
            flush.fill_(i & 0xff)
        if i >= 3:
            ev[i - 3][0].record()
        fwd(q, kc[:, :mx], vc[:, :mx], kn, vn, cache_seqlens=sl, cache_batch_idx=idx, causal=True)
        if i >= 3:
            ev[i - 3][1].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
    nbytes = kv_bytes + 2 * 2 * B * Hq * D + 2 * 2 * B * Hkv * D
    hbm, _, _ = peaks()
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": f"decode B{B} Hq{Hq} Hkv{Hkv} D128 ctx{ctx}{' ragged U[ctx/2,ctx)' if args.ragged else ''} "
                                  f"{args.dtype}, 1 layer-call incl. append", "impl": args.impl,
                      "ms_per_layer_call_median": round(ms, 4), "bytes": nbytes, "gbps": round(gbs, 1),
                      "frac_of_measured_copy_peak": round(gbs / hbm, 4), "frac_of_8TBps": round(gbs / 8000, 4),
                      "l2_flush_between_calls": flush is not None}))
    va.cleanup()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["prefill", "pod", "alloc", "decode"])
    ap.add_argument("--impl", default="ours", choices=["ours", "fa", "fi"])
    ap.add_argument("--chunk", type=int, default=2048)
    ap.add_argument("--ctx", type=int, default=None, help="context length (prefill: 131072, decode: 32768)")
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--prefills", type=int, default=8)
    ap.add_argument("--prefill-len", type=int, default=16384)
    ap.add_argument("--alloc-steps", type=int, default=48)
    ap.add_argument("--page-kb", type=int, default=2048)
    ap.add_argument("--prefill-chunk", type=int, default=0)
    ap.add_argument("--decodes", type=int, default=56)
    ap.add_argument("--decode-len", type=int, default=4096)
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--hq", type=int, default=32)
    ap.add_argument("--hkv", type=int, default=8)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--calls", type=int, default=20)
    a = ap.parse_args()
    if a.ctx is None:
        a.ctx = 32768 if a.what == "decode" else 131072
    {"prefill": prefill, "pod": pod, "alloc": alloc, "decode": decode}[a.what](a)
