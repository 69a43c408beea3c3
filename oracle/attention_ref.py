"""CPU restatement of the attention arithmetic on the vAttention hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by vattention_b200/.

The reference repo holds no attention arithmetic of its own on the fa_vattn / fi_vattn
paths: it calls two un-vendored third-party packages,
    flash-attn == 2.5.9.post1   (sarathi-lean/requirements.txt:22)
    flashinfer == 0.0.6         (sarathi-lean/requirements.txt:18)
from vattention_flashattention_wrapper.py:159-166,194-205 and
vattention_flashinfer_wrapper.py:151-158.  What is restated here is FlashAttention-2's
published forward algorithm as it appears in the FA-2.6.1 fork the reference DOES vendor
for POD (pod_attn/pod_attn/):
    flash_api.cpp:1291-1580   mha_fwd_kvcache argument semantics (append, cache_seqlens,
                              cache_batch_idx, GQA, seqlen_q==1 => causal is a no-op)
    block_info.h:11-44        actual_seqlen_k = cache_len + seqlen_knew
    mask.h:172                causal: key j visible to query i iff j <= i + Lk - Sq
    softmax.h:66-160          softmax in fp32, fully masked row -> output 0
    flash_fwd_kernel.h:685-790  new k/v rows are written to the cache BEFORE attending
Everything is computed in float32 with explicit loops over batch entries and matmuls per
head group: a plain O(Sq*Sk) reference, no tiling, no online softmax.

Parity pinning: the reference has no golden vectors for attention; its only
assertions are POD-vs-FA2 allclose(atol=1e-3) in benchmark scripts.  tests/golden/attn_*.pt
hold outputs of flash_attn 2.8.3's flash_attn_with_kvcache (the un-vendored dependency,
newer pin) generated on a B200 by oracle/gen_attn_golden.py; tests/test_oracle_golden.py
checks this file against them.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def cache_flat_ref(key: torch.Tensor, value: torch.Tensor, k_cache: torch.Tensor,
                   v_cache: torch.Tensor) -> None:
    """sarathi-lean/csrc/cache_kernels.cu:482-520: k_cache[t, :, :] = key[t, :, :]."""
    n = key.shape[0]
    k_cache[:n].copy_(key)
    v_cache[:n].copy_(value)


def rotary_ref(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, positions: torch.Tensor,
               interleaved: bool) -> torch.Tensor:
    """x [S, H, D] rotated at `positions` [S] (rotary.h / flash_fwd_kernel.h:684-830):
    interleaved pairs dims (2j, 2j+1), otherwise (j, j + rotary_dim/2); dims >= rotary_dim pass
    through; x0' = x0 cos - x1 sin, x1' = x0 sin + x1 cos in fp32, rounded once to x.dtype."""
    rd = 2 * cos.shape[1]
    c = cos[positions.long()].float().unsqueeze(1)          # [S, 1, rd/2]
    s = sin[positions.long()].float().unsqueeze(1)
    xf = x.float()
    out = xf.clone()
    if interleaved:
        x0, x1 = xf[..., 0:rd:2], xf[..., 1:rd:2]
        out[..., 0:rd:2] = x0 * c - x1 * s
        out[..., 1:rd:2] = x0 * s + x1 * c
    else:
        x0, x1 = xf[..., :rd // 2], xf[..., rd // 2:rd]
        out[..., :rd // 2] = x0 * c - x1 * s
        out[..., rd // 2:rd] = x0 * s + x1 * c
    return out.to(x.dtype)


def attn_with_kvcache_ref(
    q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
    k: Optional[torch.Tensor] = None, v: Optional[torch.Tensor] = None,
    cache_seqlens: Optional[torch.Tensor] = None, cache_batch_idx: Optional[torch.Tensor] = None,
    softmax_scale: Optional[float] = None, causal: bool = False, update_cache: bool = True,
    return_lse: bool = False, rotary_cos: Optional[torch.Tensor] = None,
    rotary_sin: Optional[torch.Tensor] = None, rotary_interleaved: bool = True,
):
    """flash_attn_with_kvcache on CPU in fp32.  q [B,Sq,Hq,D]; caches [Bc,Sk,Hkv,D] (updated in
    place when k/v are given and update_cache); returns out [B,Sq,Hq,D] in q.dtype."""
    B, Sq, Hq, D = q.shape
    Bc, Sk, Hkv, _ = k_cache.shape
    g = Hq // Hkv
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    out = torch.zeros((B, Sq, Hq, D), dtype=torch.float32)
    lse = torch.full((B, Hq, Sq), float("inf"), dtype=torch.float32)
    for b in range(B):
        slot = int(cache_batch_idx[b]) if cache_batch_idx is not None else b
        L0 = int(cache_seqlens[b]) if cache_seqlens is not None else Sk
        qb = q[b]
        if k is not None:
            n_new = k.shape[1]
            kb = k[b]
            if rotary_cos is not None:
                # flash_fwd_kernel.h:693-760 (new keys at L0 + t), :796-804 (queries at L0 + i when
                # causal, all at L0 otherwise)
                kb = rotary_ref(kb, rotary_cos, rotary_sin, L0 + torch.arange(n_new), rotary_interleaved)
                qpos = L0 + (torch.arange(Sq) if causal else torch.zeros(Sq, dtype=torch.long))
                qb = rotary_ref(qb, rotary_cos, rotary_sin, qpos, rotary_interleaved)
            if update_cache:
                k_cache[slot, L0:L0 + n_new] = kb         # flash_fwd_kernel.h:685-790
                v_cache[slot, L0:L0 + n_new] = v[b]
            kk = torch.cat([k_cache[slot, :L0], kb], dim=0).float()
            vv = torch.cat([v_cache[slot, :L0], v[b]], dim=0).float()
            Lk = L0 + n_new                                # block_info.h:37-41
        else:
            Lk = L0
            kk = k_cache[slot, :Lk].float()
            vv = v_cache[slot, :Lk].float()
        if Lk == 0:
            continue
        qq = qb.float()                                    # [Sq, Hq, D]
        # GQA: q head h reads kv head h // g
        kk = kk.repeat_interleave(g, dim=1)                # [Lk, Hq, D]
        vv = vv.repeat_interleave(g, dim=1)
        s = torch.einsum("ihd,jhd->hij", qq, kk) * scale   # [Hq, Sq, Lk]
        if causal:
            i = torch.arange(Sq).view(-1, 1)
            j = torch.arange(Lk).view(1, -1)
            s = s.masked_fill(~(j <= i + (Lk - Sq)), float("-inf"))  # mask.h:172
        m = s.max(dim=-1, keepdim=True).values
        dead = torch.isinf(m) & (m < 0)                    # fully masked rows (softmax.h:76-78)
        m = torch.where(dead, torch.zeros_like(m), m)
        p = torch.exp(s - m)
        denom = p.sum(dim=-1, keepdim=True)
        o = torch.einsum("hij,jhd->ihd", p / torch.where(denom == 0, torch.ones_like(denom), denom), vv)
        o = torch.where(dead.permute(1, 0, 2).expand(Sq, Hq, 1), torch.zeros_like(o), o)
        out[b] = o
        l = (m + torch.log(denom)).squeeze(-1)
        lse[b] = torch.where(dead.squeeze(-1), torch.full_like(l, float("inf")), l)
    out = out.to(q.dtype)
    return (out, lse) if return_lse else out


def single_prefill_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True,
                       sm_scale: Optional[float] = None) -> torch.Tensor:
    """flashinfer.single_prefill_with_kv_cache(q[c,Hq,D], k[n,Hkv,D], v, causal): one request,
    bottom-right aligned causal mask, scale 1/sqrt(D) (vattention_flashinfer_wrapper.py:151-158)."""
    return attn_with_kvcache_ref(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), causal=causal,
                                 softmax_scale=sm_scale).squeeze(0)


def pod_ref(q_p, k_cache_p, v_cache_p, q_d, k_cache_d, v_cache_d, k=None, v=None,
            cache_seqlens_p=None, cache_seqlens_d=None, cache_batch_idx=None,
            softmax_scale=None, causal=True) -> Tuple[torch.Tensor, torch.Tensor]:
    """POD's contract is 'each output equals the separate FA call' (pod_attn/tests/attn_sweep.py:
    82-97 asserts exactly that), so the oracle is two independent calls."""
    out_p = attn_with_kvcache_ref(q_p, k_cache_p, v_cache_p, cache_seqlens=cache_seqlens_p,
                                  softmax_scale=softmax_scale, causal=causal)
    out_d = attn_with_kvcache_ref(q_d, k_cache_d, v_cache_d, k=k, v=v,
                                  cache_seqlens=cache_seqlens_d, cache_batch_idx=cache_batch_idx,
                                  softmax_scale=softmax_scale, causal=causal)
    return out_p, out_d


def sdpa_decode_cpu(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float) -> torch.Tensor:
    """BASELINE.json configs[0]: '1-seq 8-head x 128-dim 1K-ctx decode attn via torch SDPA on CPU'.
    q [B,1,Hq,D], k/v [B,L,Hkv,D] already holding the appended token.  Used as the timed CPU
    baseline (bench.py) -- torch's own fused CPU kernel, not the loop above; GQA through SDPA's
    enable_gqa so that no 4x copy of K/V is materialised inside the timed region."""
    qq = q.transpose(1, 2)                                  # [B,Hq,1,D]
    kk, vv = k.transpose(1, 2), v.transpose(1, 2)           # [B,Hkv,L,D] views: no copy of K/V
    o = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=False, scale=scale,
                                                         enable_gqa=True)
    return o.transpose(1, 2)
