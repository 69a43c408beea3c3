"""CPU: the FMA-pipe exp2 used by the opt-in prefill variant (attn_tc_work.cuh::poly_exp2), emulated in
float32 with the constants read from the source: relative error against 2^x over the range the softmax
feeds it ([-126, 8]: exponent arguments are <= the lazy-rescale threshold above the reference max)."""
import re
from pathlib import Path

import numpy as np

SRC = Path(__file__).resolve().parent.parent / "vattention_b200" / "csrc" / "attn_tc_work.cuh"


def constants():
    body = SRC.read_text().split("__device__ __forceinline__ float poly_exp2(float x) {")[1].split("}")[0]
    vals = [np.float32(v) for v in re.findall(r"([0-9]+\.[0-9]*)f", body)]
    # -126 clamp, magic (twice), c3, c2, c1, c0
    assert vals[0] == np.float32(126.0) and vals[1] == vals[2] == np.float32(12582912.0)
    return vals[3:7]


def poly_exp2(x, c3, c2, c1, c0):
    x = np.maximum(x.astype(np.float32), np.float32(-126.0))
    magic = np.float32(12582912.0)
    t = (x + magic).astype(np.float32)
    f = (x - (t - magic).astype(np.float32)).astype(np.float32)
    p = (c3 * f + c2).astype(np.float32)
    p = (p * f + c1).astype(np.float32)
    p = (p * f + c0).astype(np.float32)
    return (p.view(np.int32) + (t.view(np.int32) << np.int32(23))).view(np.float32)


def test_poly_exp2_accuracy_and_range():
    c3, c2, c1, c0 = constants()
    xs = np.concatenate([np.linspace(-125.0, 8.0, 1_000_001), np.arange(-125, 9) + 0.5, np.arange(-125, 9) - 0.5,
                         np.arange(-125, 9)]).astype(np.float32)
    got = poly_exp2(xs, c3, c2, c1, c0).astype(np.float64)
    ref = 2.0 ** xs.astype(np.float64)
    rel = np.abs(got - ref) / ref
    assert rel.max() < 8e-5, rel.max()            # a quarter of fp16's 2^-11 rounding step
    # monotone where it matters (P must not reorder keys by more than its own error) and clamped, never NaN
    lo = poly_exp2(np.array([-1e30, -np.inf, -200.0], dtype=np.float32), c3, c2, c1, c0)
    assert np.all(np.isfinite(lo)) and np.all(lo < 2e-38) and np.all(lo > 0)
