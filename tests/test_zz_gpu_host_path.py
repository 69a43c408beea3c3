"""GPU: the host-buffer entry points of the C ABI (what bench.py's `e2e` leg times) against the oracle.

vattn_fwd_kvcache_host / _host_async take q, k_new, v_new, cache_seqlens, cache_batch_idx and out in HOST
memory, the caches on the device.  (File name sorts last on purpose: these entry points were exercised by
the bench long before they had a parity test.)  The pipelined variant (copies on their own streams) is what the
bench's e2e leg uses: measured 1679 vs 1606 tokens/s for the in-line copies.
"""
import pytest
import torch

from oracle import attention_ref as ref
from vattention_b200 import attention as att

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def make(seed, B=5, Hq=32, Hkv=8, D=128, Sk=3000, slots=7, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, 1, Hq, D, generator=g).to(dtype)
    kc = torch.randn(slots, Sk, Hkv, D, generator=g).to(dtype)
    vc = torch.randn(slots, Sk, Hkv, D, generator=g).to(dtype)
    kn = torch.randn(B, 1, Hkv, D, generator=g).to(dtype)
    vn = torch.randn(B, 1, Hkv, D, generator=g).to(dtype)
    lens = torch.randint(1, Sk - 1, (B,), generator=g).int()
    lens[0] = Sk - 1
    idx = torch.randperm(slots, generator=g)[:B].int()
    return q, kc, vc, kn, vn, lens, idx


def check(out, want):
    want = want.float()
    tol = 3e-3 * want.abs().max() + 2.0 ** -7 * want.abs() + 1e-6
    assert torch.all((out.float() - want).abs() <= tol), (out.float() - want).abs().max()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("wait", [True, False])
def test_host_buffers_match_oracle(wait):
    q, kc, vc, kn, vn, lens, idx = make(1)
    kc_ref, vc_ref = kc.clone(), vc.clone()
    want = ref.attn_with_kvcache_ref(q, kc_ref, vc_ref, kn, vn, lens, idx, None, True)
    kd, vd = kc.to(DEV), vc.to(DEV)
    pin = lambda t: t.contiguous().pin_memory()
    out = torch.empty_like(q).pin_memory()
    att.flash_attn_with_kvcache_host(pin(q), kd, vd, pin(kn), pin(vn), pin(lens), pin(idx), out, causal=True,
                                     wait=wait)
    torch.cuda.synchronize()
    check(out, want)
    assert torch.equal(kd.cpu(), kc_ref) and torch.equal(vd.cpu(), vc_ref)   # append landed on the device


@pytest.mark.timeout(120)
def test_pipelined_host_calls_match_oracle():
    """Eight back-to-back pipelined calls with different inputs (both staging slots reused several times),
    one join, one synchronise; every output and the final cache state must match."""
    cases = [make(10 + i) for i in range(8)]
    kd, vd = cases[0][1].to(DEV), cases[0][2].to(DEV)
    kc_ref, vc_ref = cases[0][1].clone(), cases[0][2].clone()
    pin = lambda t: t.contiguous().pin_memory()
    outs, wants, keep = [], [], []
    for q, _, _, kn, vn, lens, idx in cases:
        wants.append(ref.attn_with_kvcache_ref(q, kc_ref, vc_ref, kn, vn, lens, idx, None, True))
        host = [pin(t) for t in (q, kn, vn, lens, idx)]
        keep.append(host)
        out = torch.empty_like(q).pin_memory()
        outs.append(out)
        att.flash_attn_with_kvcache_host(host[0], kd, vd, host[1], host[2], host[3], host[4], out, causal=True,
                                         wait=False, pipelined=True)
    att.host_pipeline_join(DEV)
    torch.cuda.synchronize()
    for out, want in zip(outs, wants):
        check(out, want)
    assert torch.equal(kd.cpu(), kc_ref) and torch.equal(vd.cpu(), vc_ref)
