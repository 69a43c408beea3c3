#!/bin/bash
# Round 2, second eight-GPU session: the headline workload at N = 8 with the per-N crossing window
# (the first session's step time tracked the mapper's background pass: 16 pages / step / rank at
# ~730 us each with eight processes in the driver at once), then configs[4] at 64K again.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_n8_window.log) 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "=== decode32k, N = 8 (fused + graph)"; timeout 150 $TR --master-port 29651 bench.py --gpus 8 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp8_fused_window.json | cut -c1-300
echo "=== tp70b ctx 65536, N = 8"; timeout 100 $TR --master-port 29654 bench.py --gpus 8 --workload tp70b --ctx 65536 --no-e2e 2>&1 | grep '^{' | tee gpurun_out/r2_bench_tp70b_n8_64k_window.json | cut -c1-300
echo "=== done"
