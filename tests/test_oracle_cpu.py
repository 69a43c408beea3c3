"""Oracle self-checks on CPU: BASELINE configs[0] (1 seq, 8 heads x 128, 1K ctx decode via
torch SDPA) and algebraic properties of the restated algorithm."""
import math

import torch

from oracle import attention_ref as ref


def test_config1_decode_equals_torch_sdpa():
    # SURVEY 8(d) config 1
    torch.manual_seed(0)
    q = torch.randn(1, 1, 8, 128)
    k = torch.randn(1, 1024, 8, 128)
    v = torch.randn(1, 1024, 8, 128)
    kn, vn = k[:, 1023:].clone(), v[:, 1023:].clone()
    kc, vc = k.clone(), v.clone()
    kc[:, 1023:] = 0
    vc[:, 1023:] = 0
    scale = 128 ** -0.5
    out = ref.attn_with_kvcache_ref(q, kc, vc, kn, vn, torch.tensor([1023], dtype=torch.int32),
                                    softmax_scale=scale, causal=True)
    want = ref.sdpa_decode_cpu(q, k, v, scale)
    assert torch.allclose(out, want, atol=1e-5, rtol=1e-5)
    assert torch.equal(kc, k) and torch.equal(vc, v)  # append wrote row 1023
    for dt in (torch.bfloat16,):
        out16 = ref.attn_with_kvcache_ref(q.to(dt), k.to(dt), v.to(dt), softmax_scale=scale)
        assert torch.allclose(out16.float(), want, atol=2e-2)


def test_causal_is_bottom_right_aligned():
    torch.manual_seed(1)
    q = torch.randn(1, 4, 2, 64)
    k = torch.randn(1, 10, 2, 64)
    v = torch.randn(1, 10, 2, 64)
    out = ref.attn_with_kvcache_ref(q, k, v, causal=True)
    # last query row sees all 10 keys, first sees 7 (mask.h:172: j <= i + Lk - Sq)
    for i, n in ((3, 10), (0, 7)):
        s = torch.einsum("hd,jhd->hj", q[0, i], k[0, :n]) / math.sqrt(64)
        want = torch.einsum("hj,jhd->hd", torch.softmax(s, -1), v[0, :n])
        assert torch.allclose(out[0, i], want, atol=1e-5)


def test_gqa_and_batch_idx_and_lse():
    torch.manual_seed(2)
    q = torch.randn(2, 1, 6, 64)
    k = torch.randn(3, 20, 2, 64)
    v = torch.randn(3, 20, 2, 64)
    lens = torch.tensor([20, 7], dtype=torch.int32)
    idx = torch.tensor([2, 0], dtype=torch.int32)
    out, lse = ref.attn_with_kvcache_ref(q, k, v, cache_seqlens=lens, cache_batch_idx=idx, return_lse=True)
    h = 4  # q head 4 -> kv head 1
    s = (q[1, 0, h] @ k[0, :7, 1].T) / 8.0
    assert torch.allclose(out[1, 0, h], torch.softmax(s, -1) @ v[0, :7, 1], atol=1e-5)
    assert torch.allclose(lse[1, h, 0], torch.logsumexp(s, -1), atol=1e-5)


def test_fully_masked_rows_zero_and_empty_cache():
    q = torch.randn(1, 5, 2, 64)
    k = torch.randn(1, 8, 2, 64)
    out = ref.attn_with_kvcache_ref(q, k, k.clone(), cache_seqlens=torch.tensor([2], dtype=torch.int32), causal=True)
    assert torch.all(out[0, :3] == 0) and not torch.all(out[0, 3] == 0)
    out = ref.attn_with_kvcache_ref(q, k, k.clone(), cache_seqlens=torch.tensor([0], dtype=torch.int32))
    assert torch.all(out == 0)


def test_rotary_restatement_properties():
    """rotary_ref (flash_fwd_kernel.h:684-830): rotations preserve norms, position 0 of a standard
    frequency table is the identity, and <R(p) q, R(p') k> depends on p - p' only -- for both pairings;
    dims beyond rotary_dim pass through; queries of a non-causal call all sit at cache_seqlens."""
    import torch
    from oracle.attention_ref import attn_with_kvcache_ref, rotary_ref
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 3, 64, generator=g)
    inv = 1.0 / (10000 ** (torch.arange(0, 24) / 24))
    ang = torch.arange(100)[:, None] * inv[None]
    cos, sin = torch.cos(ang), torch.sin(ang)           # rotary_dim 48 of 64
    for il in (True, False):
        y = rotary_ref(x, cos, sin, torch.tensor([0, 3, 7, 50, 99]), il)
        assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-4)
        assert torch.equal(y[0], x[0]) and torch.equal(y[..., 48:], x[..., 48:])
        q, k = torch.randn(1, 1, 64, generator=g), torch.randn(1, 1, 64, generator=g)
        a = (rotary_ref(q, cos, sin, torch.tensor([10]), il) * rotary_ref(k, cos, sin, torch.tensor([4]), il)).sum()
        b = (rotary_ref(q, cos, sin, torch.tensor([56]), il) * rotary_ref(k, cos, sin, torch.tensor([50]), il)).sum()
        assert abs(a - b) < 1e-3
    # end to end: rotating by hand then calling without rotary == calling with rotary
    q = torch.randn(2, 3, 4, 64, generator=g)
    kc, vc = torch.randn(2, 40, 2, 64, generator=g), torch.randn(2, 40, 2, 64, generator=g)
    kn, vn = torch.randn(2, 3, 2, 64, generator=g), torch.randn(2, 3, 2, 64, generator=g)
    lens = torch.tensor([20, 37], dtype=torch.int32)
    for causal in (True, False):
        k1, v1, k2, v2 = kc.clone(), vc.clone(), kc.clone(), vc.clone()
        got = attn_with_kvcache_ref(q, k1, v1, kn, vn, lens, None, None, causal, rotary_cos=cos, rotary_sin=sin,
                                    rotary_interleaved=False)
        qr = torch.stack([rotary_ref(q[b], cos, sin, int(lens[b]) + (torch.arange(3) if causal else torch.zeros(3, dtype=torch.long)), False)
                          for b in range(2)])
        kr = torch.stack([rotary_ref(kn[b], cos, sin, int(lens[b]) + torch.arange(3), False) for b in range(2)])
        want = attn_with_kvcache_ref(qr, k2, v2, kr, vn, lens, None, None, causal)
        assert torch.equal(got, want) and torch.equal(k1, k2) and torch.equal(v1, v2)
