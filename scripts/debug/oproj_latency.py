"""single GPU: latency of the fused o_proj kernel with world = 1 (GEMM + self push/flag/reduce) vs torch.matmul."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from vattention_b200.tp import FusedOProjAllReduce
dev = torch.device("cuda", 0)

def bench(fn, n=300):
    for _ in range(30):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 2)

for K in (2048, 512):
    w = (torch.randn(4096, K, device=dev) * 0.05).bfloat16()
    wt = w.t().contiguous()
    x = torch.randn(64, K, device=dev).bfloat16()
    op = FusedOProjAllReduce(w, 128, local_only=True)
    print({"K": K, "env": {k: v for k, v in os.environ.items() if k.startswith("VATTN_OPROJ")},
           "fused_world1_us": bench(lambda: op(x)), "matmul_us": bench(lambda: torch.matmul(x, wt))}, flush=True)
