#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/final_n4.log) 2>&1
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== bench default (nccl + graph, with e2e)"; timeout 240 $TR --master-port 29621 bench.py --gpus $N 2>&1 | grep '^{' | tee gpurun_out/bench_tp4_nccl_graph.json
echo "=== bench fused + graph"; timeout 200 $TR --master-port 29622 bench.py --gpus $N --no-e2e --tp-collective fused 2>&1 | grep '^{' | tee gpurun_out/bench_tp4_fused_graph.json
echo "=== done"
