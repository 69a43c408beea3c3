#!/bin/bash
# First GPU pass: parity tests, golden-vector generation from the reference, smoke, short bench.
set -u
mkdir -p gpurun_out/golden
exec > >(tee gpurun_out/round1.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv
nproc
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -x --deselect tests/test_oracle_golden.py 2>&1 | tail -60
echo "=== golden alloc"; timeout 900 python oracle/gen_alloc_golden.py gpurun_out/golden
echo "=== golden attn"; timeout 600 python oracle/gen_attn_golden.py gpurun_out/golden
echo "=== bench ours"; timeout 900 python bench.py --steps 3 --warmup 3 | tee gpurun_out/bench_ours.json
echo "=== bench fa_vattn"; timeout 600 python bench.py --impl fa_vattn --steps 3 --warmup 2 | tee gpurun_out/bench_fa.json
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_ref.json
echo "=== done"
