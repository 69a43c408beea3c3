#!/bin/bash
set -u
mkdir -p gpurun_out/prof
exec > >(tee gpurun_out/round4.log) 2>&1
echo "=== full gpu suite"; timeout 1800 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
echo "=== pod fused"; timeout 600 python scripts/bench_extra.py pod --impl ours
echo "=== pod hybrid (chunk 1024 @16K prefix + decode)"; timeout 600 python scripts/bench_extra.py pod --impl ours --prefills 1 --prefill-len 16384
echo "=== bench ours"; timeout 900 python bench.py | tee gpurun_out/bench_ours4.json
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1
echo "=== done"
