#!/bin/bash
set -u
exec > >(tee gpurun_out/round5.log) 2>&1
echo "=== pod debug"; timeout 300 python scripts/debug/pod_debug.py 2>&1 | tail -20 | cut -c1-200
echo "=== attention tests"; timeout 900 python -m pytest tests/test_gpu_attention.py -q --timeout 300 2>&1 | tail -8
echo "=== prefill ours"; for c in 512 2048 8192; do timeout 600 python scripts/bench_extra.py prefill --chunk $c --impl ours; done
echo "=== pod"; timeout 600 python scripts/bench_extra.py pod --impl ours; timeout 600 python scripts/bench_extra.py pod --impl ours --prefills 1 --prefill-len 16384
echo "=== done"
