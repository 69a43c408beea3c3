#!/bin/bash
# 2 GPUs: fused o_proj + all-reduce parity across ranks, then the TP bench arms on the same box
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/round7_n2.log) 2>&1
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== fused parity + latency (world $N)"; timeout 300 $TR --master-port 29601 scripts/debug/tp_fused_test.py
echo "=== bench fused + graph"; timeout 400 $TR --master-port 29602 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --tp-collective fused | tee gpurun_out/bench_tp${N}_fused_graph.json
echo "=== bench fused eager"; timeout 400 $TR --master-port 29603 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --tp-collective fused --no-tp-graph | tee gpurun_out/bench_tp${N}_fused_eager.json
echo "=== bench peer eager (previous default)"; timeout 400 $TR --master-port 29604 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --tp-collective peer | tee gpurun_out/bench_tp${N}_peer_eager.json
echo "=== bench nccl + graph"; timeout 400 $TR --master-port 29605 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --tp-collective nccl | tee gpurun_out/bench_tp${N}_nccl_graph.json
echo "=== done"
