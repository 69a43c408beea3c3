#!/bin/bash
set -u
exec > >(tee gpurun_out/round6.log) 2>&1
echo "=== attention tests"; timeout 900 python -m pytest tests/test_gpu_attention.py -q --timeout 300 2>&1 | tail -12
echo "=== prefill ours (2 row blocks)"; for c in 512 2048 8192; do timeout 600 python scripts/bench_extra.py prefill --chunk $c --impl ours; done
echo "=== prefill ours (single block, for comparison)"; VATTN_PREFILL_SINGLE=1 timeout 600 python scripts/bench_extra.py prefill --chunk 2048 --impl ours
echo "=== pod"; timeout 600 python scripts/bench_extra.py pod --impl ours; timeout 600 python scripts/bench_extra.py pod --impl ours --prefills 1 --prefill-len 16384
echo "=== done"
