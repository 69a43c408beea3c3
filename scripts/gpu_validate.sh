#!/bin/bash
# One-GPU validation pass used during development (run under gpurun from the repo root):
#   parity suite, smoke, headline bench, secondary workloads.  Output -> gpurun_out/validate.log
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/validate.log) 2>&1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 90 2>&1 | tail -15
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
echo "=== bench (ours)"; timeout 900 python bench.py | tee gpurun_out/bench_ours.json
if [ "${1:-}" = "full" ]; then
  echo "=== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1
  echo "=== bench (flash_attn library over vAttention tensors)"; timeout 600 python bench.py --impl fa_vattn --steps 3 --warmup 2
  for c in 512 2048 8192; do timeout 600 python scripts/bench_extra.py prefill --chunk $c; done
  timeout 600 python scripts/bench_extra.py pod
  timeout 600 python scripts/bench_extra.py alloc
fi
echo "=== done"
