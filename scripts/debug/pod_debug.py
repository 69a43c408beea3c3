"""Locate mismatches of the fused POD kernel against the two separate calls (bitwise)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from vattention_b200 import attention as att

DEV = "cuda:0"
for dtype in (torch.float16, torch.bfloat16, torch.float16):
    g = torch.Generator().manual_seed(77)
    Hq, Hkv, D = 8, 2, 128
    Bp, Sq, Sk = 2, 700, 1500
    q_p = torch.randn(Bp, Sq, Hq, D, generator=g).to(dtype).to(DEV)
    kc_p = torch.randn(Bp, Sk, Hkv, D, generator=g).to(dtype).to(DEV)
    vc_p = torch.randn(Bp, Sk, Hkv, D, generator=g).to(dtype).to(DEV)
    lens_p = torch.tensor([1500, 901], dtype=torch.int32).to(DEV)
    Bd, Sd = 37, 3000
    q_d = torch.randn(Bd, 1, Hq, D, generator=g).to(dtype).to(DEV)
    kc_d = torch.randn(Bd + 3, Sd, Hkv, D, generator=g).to(dtype).to(DEV)
    vc_d = torch.randn(Bd + 3, Sd, Hkv, D, generator=g).to(dtype).to(DEV)
    lens_d = torch.randint(1, Sd, (Bd,), generator=g).int().to(DEV)
    idx = torch.randperm(Bd + 3, generator=g)[:Bd].int().to(DEV)
    sep_p = att.flash_attn_with_kvcache(q_p, kc_p, vc_p, cache_seqlens=lens_p, causal=True)
    sep_d = att.flash_attn_with_kvcache(q_d, kc_d, vc_d, cache_seqlens=lens_d, cache_batch_idx=idx, causal=True)
    for it in range(6):
        out_p, out_d = att.true_fused_attn_with_kvcache(q_p, kc_p, vc_p, q_d, kc_d, vc_d, None, None, causal=True,
                                                        cache_seqlens_p=lens_p, cache_seqlens_d=lens_d,
                                                        cache_batch_idx=idx)
        torch.cuda.synchronize()
        bad_p = (out_p != sep_p).any(dim=-1)            # [Bp, Sq, Hq]
        bad_d = (out_d != sep_d).any(dim=-1)[:, 0]      # [Bd, Hq]
        msg = f"{dtype} iter {it}: prefill bad rows {int(bad_p.sum())}, decode bad (b,h) {int(bad_d.sum())}"
        if bad_p.any():
            b, i, h = torch.nonzero(bad_p, as_tuple=True)
            tiles = sorted(set((int(x), int(y) // 128, int(z)) for x, y, z in zip(b, i, h)))
            msg += f" | prefill (b, mtile, h): {tiles[:12]} rows in tile e.g. {sorted(set(int(y) % 128 for y in i))[:8]}"
            e = (out_p.float() - sep_p.float()).abs()
            msg += f" maxerr {e.max().item():.3e} nan {int(torch.isnan(out_p.float()).sum())}"
        if bad_d.any():
            b, h = torch.nonzero(bad_d, as_tuple=True)
            msg += f" | decode (b,h): {list(zip(b.tolist(), h.tolist()))[:12]} lens {lens_d[b[:6]].tolist()}"
            e = (out_d.float() - sep_d.float()).abs()
            msg += f" maxerr {e.max().item():.3e}"
        print(msg, flush=True)
