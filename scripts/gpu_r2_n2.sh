#!/bin/bash
# Round 2, two-GPU session: fused o_proj + all-reduce parity across ranks, collective arms, bench at N = 2.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_n2.log) 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== pytest multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_tp_peer.py tests/test_gpu_oproj.py -q --timeout 300 --tb=short 2>&1 | tail -6
echo "=== fused GEMM + all-reduce parity / latency at world 2"
timeout 300 $TR --master-port 29631 scripts/debug/tp_fused_test.py 2>&1 | grep -v "^\*\|OMP\|^$" | tail -6
echo "=== bench default (fused + graph)"; timeout 400 $TR --master-port 29632 bench.py --gpus 2 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp2_fused.json | cut -c1-1500
echo "=== bench nccl + graph"; timeout 400 $TR --master-port 29633 bench.py --gpus 2 --no-e2e --tp-collective nccl 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp2_nccl.json | cut -c1-600
echo "=== tp70b at N=2 (ctx 32768)"; timeout 400 $TR --master-port 29634 bench.py --gpus 2 --workload tp70b --no-e2e 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp70b_n2.json | cut -c1-800
echo "=== done"
