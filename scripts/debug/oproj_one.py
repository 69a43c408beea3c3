"""one call of the fused o_proj kernel (world 1) -- for compute-sanitizer"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from vattention_b200.tp import FusedOProjAllReduce
dev = torch.device("cuda", 0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = (torch.randn(4096, K, device=dev) * 0.05).bfloat16()
x = torch.randn(64, K, device=dev).bfloat16()
op = FusedOProjAllReduce(w, 128, local_only=True)
y = op(x).clone()
torch.cuda.synchronize()
ref = (x.float() @ w.float().t())
print("max err", (y.float() - ref).abs().max().item(), "failed", op.failed())
