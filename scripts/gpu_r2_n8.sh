#!/bin/bash
# Round 2, eight-GPU session: the headline workload and configs[4] (Llama-3-70B TP = 8) at N = 8.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_n8.log) 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "=== decode32k, N = 8 (fused + graph)"; timeout 300 $TR --master-port 29641 bench.py --gpus 8 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp8_fused.json | cut -c1-300
echo "=== tp70b ctx 32768, N = 8"; timeout 240 $TR --master-port 29642 bench.py --gpus 8 --workload tp70b --ctx 32768 --no-e2e 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp70b_n8_32k.json | cut -c1-300
echo "=== tp70b ctx 131072, N = 8"; timeout 240 $TR --master-port 29643 bench.py --gpus 8 --workload tp70b --ctx 131072 --no-e2e 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp70b_n8_128k.json | cut -c1-300
echo "=== tp70b ctx 65536, N = 8"; timeout 240 $TR --master-port 29644 bench.py --gpus 8 --workload tp70b --ctx 65536 --no-e2e 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp70b_n8_64k.json | cut -c1-300
echo "=== done"
