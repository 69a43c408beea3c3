"""torchrun --nproc-per-node N scripts/debug/tp_peer_test.py : PeerAllReduce vs NCCL (parity + latency)."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from vattention_b200.tp import PeerAllReduce

T, H = 64, 4096
ar = PeerAllReduce(H, 256, torch.bfloat16, dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(50):
    tokens = [64, 1, 17, 256][it % 4]
    x = torch.randn(tokens, H, device=dev, generator=g).bfloat16()
    ar.partial_buffer(tokens).copy_(x)
    got = ar.reduce(tokens).clone()
    # exact reference: gather every rank's bf16 partial, sum in fp32, round once (what the kernel does)
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    acc = torch.zeros_like(parts[0], dtype=torch.float32)
    for p in parts:                      # same order as the kernel: rank 0 .. world-1, fp32
        acc += p.float()
    want = acc.bfloat16()
    worst = max(worst, (got.float() - want.float()).abs().max().item())
    assert torch.equal(got, want), f"rank {rank} iter {it}: mismatch {worst}"
if rank == 0:
    print(f"PeerAllReduce == fp32-sum-of-partials bitwise over 50 calls (world {world})", flush=True)

def bench(fn, n=200):
    for _ in range(20):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

x = torch.randn(T, H, device=dev).bfloat16()
w = (torch.randn(512, H, device=dev) * 0.02).bfloat16()
a = torch.randn(T, 512, device=dev).bfloat16()
y = torch.empty_like(x)
def nccl_only():
    dist.all_reduce(x)
def peer_only():
    ar.reduce(T)
def gemm_nccl():
    p = a @ w
    dist.all_reduce(p)
def gemm_peer():
    torch.matmul(a, w, out=ar.partial_buffer(T))
    ar.reduce(T)
res = {k: round(bench(f), 2) for k, f in (("nccl_allreduce_us", nccl_only), ("peer_allreduce_us", peer_only),
                                             ("gemm+nccl_us", gemm_nccl), ("gemm+peer_us", gemm_peer))}
if rank == 0:
    print({"world": world, "message": "64x4096 bf16 (512 KB)", **res}, flush=True)
dist.barrier()
dist.destroy_process_group()
