#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/round8_n2.log) 2>&1
echo "=== oproj tests (world 1)"; timeout 300 python -m pytest tests/test_gpu_oproj.py -q --timeout 60 -x 2>&1 | tail -8
echo "=== fused parity + latency (world 2)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/debug/tp_fused_test.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$"
echo "=== done"
