#!/bin/bash
# Second GPU pass: tcgen05 descriptor self test, TC decode/prefill parity, reference goldens, bench.
set -u
mkdir -p gpurun_out/golden
exec > >(tee gpurun_out/round2.log) 2>&1
echo "=== umma selftest"; timeout 300 python -m pytest tests/test_gpu_umma.py -q -s --timeout 120 2>&1 | tail -15
for v in 0 1; do
  echo "=== decode TC parity, VATTN_UMMA_MN_VARIANT=$v"
  VATTN_UMMA_MN_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_attention.py -q --timeout 120 -k "decode_append and auto" 2>&1 | tail -15
done
echo "=== prefill TC parity"; timeout 600 python -m pytest tests/test_gpu_attention.py -q --timeout 120 -k "prefill_matches and auto" 2>&1 | tail -25
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40
echo "=== golden alloc"; timeout 900 python oracle/gen_alloc_golden.py gpurun_out/golden
echo "=== bench ours"; timeout 900 python bench.py --steps 4 --warmup 3 | tee gpurun_out/bench_ours2.json
echo "=== done"
