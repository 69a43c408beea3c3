#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/round9.log) 2>&1
for rot in 5 0; do for nt in 0 64 128; do
  VATTN_OPROJ_KROT=$rot VATTN_OPROJ_NTILE=$nt timeout 120 python scripts/debug/oproj_latency.py
done; done
timeout 200 python -m pytest tests/test_gpu_oproj.py -q --timeout 60 -x 2>&1 | tail -3
echo "=== done"
