"""torchrun --nproc-per-node N scripts/debug/tp_fused_test.py : FusedOProjAllReduce (GEMM + all-reduce in one
kernel over NVLink peer memory) vs GEMM + NCCL / GEMM + PeerAllReduce: parity, then latency."""
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
from vattention_b200.tp import FusedOProjAllReduce, PeerAllReduce

H, K = 4096, 4096 // world
g = torch.Generator(device=dev).manual_seed(100 + rank)
w = (torch.randn(H, K, device=dev, generator=g) * 0.05).bfloat16()      # nn.Linear layout [hidden, k_local]
op = FusedOProjAllReduce(w, 128)
for it in range(40):
    tokens = [64, 1, 17, 128][it % 4]
    x = torch.randn(tokens, K, device=dev, generator=g).bfloat16()
    got = op(x).clone()
    # exact reference: every rank's partial rounded to bf16, summed in fp32 in rank order, rounded once
    part = (x.float() @ w.float().t()).bfloat16()
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    acc = torch.zeros_like(part, dtype=torch.float32)
    for p in parts:
        acc += p.float()
    want = acc.bfloat16()
    scale = want.float().abs().max().item()
    err = (got.float() - want.float()).abs().max().item()
    # the tensor core's fp32 accumulation order differs from torch's: a partial can round to the neighbouring bf16
    assert err <= 2 * 2.0 ** -7 * scale, f"rank {rank} iter {it}: err {err} scale {scale}"
    # every rank must hold the SAME bits (replicated activations must not diverge)
    mine = got.view(torch.int16).to(torch.int32)
    ref = mine.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(mine, ref), f"rank {rank} iter {it}: ranks disagree"
assert not op.failed()
if rank == 0:
    print(f"FusedOProjAllReduce parity ok over 40 calls, identical bits on all ranks (world {world})", flush=True)


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


T = 64
a = torch.randn(T, K, device=dev).bfloat16()
wt = w.t().contiguous()
ar = PeerAllReduce(H, 128, torch.bfloat16, dev)
def gemm_nccl():
    p = a @ wt
    dist.all_reduce(p)
def gemm_peer():
    torch.matmul(a, wt, out=ar.partial_buffer(T))
    ar.reduce(T)
def fused():
    op(a)
def gemm_only():
    torch.matmul(a, wt)
res = {k: round(bench(f), 2) for k, f in (("gemm_only_us", gemm_only), ("gemm+nccl_us", gemm_nccl),
                                             ("gemm+peer_us", gemm_peer), ("fused_us", fused))}
# the same three inside a CUDA graph of 32 calls (one decode iteration's worth): launch overhead removed
graph = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    fused()
torch.cuda.current_stream().wait_stream(s)
dist.barrier(); torch.cuda.synchronize()
with torch.cuda.graph(graph):
    for _ in range(32):
        fused()
res["fused_graph32_us_per_call"] = round(bench(graph.replay, 20) / 32, 2)
assert not op.failed()
if rank == 0:
    print({"world": world, "shape": f"[{T} x {K}] . [{K} x {H}] bf16 per rank, all-reduce of 512 KB", **res}, flush=True)
dist.barrier()
dist.destroy_process_group()
