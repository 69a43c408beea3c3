#!/bin/bash
# fused o_proj + all-reduce (world 1) parity, then the decode chunk-size sweep on the small-batch shapes
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/round6.log) 2>&1
echo "=== oproj tests"; timeout 420 python -m pytest tests/test_gpu_oproj.py -q --timeout 60 -x 2>&1 | tail -15
B="timeout 200 python scripts/bench_extra.py"
echo "=== tpc sweep: TP8 per-GPU shapes B16 ctx32K"
for t in 0 2 4 6 8 12 16; do echo "tpc=$t"; VATTN_DECODE_TPC=$t $B decode --hq 8 --hkv 1 --batch 16 --ctx 32768 | tee -a gpurun_out/tpc_sweep.jsonl; done
echo "=== tpc sweep: POD decode shape B56 ctx4K"
for t in 0 4 6 8 12 16; do echo "tpc=$t"; VATTN_DECODE_TPC=$t $B decode --hq 32 --hkv 8 --batch 56 --ctx 4096 | tee -a gpurun_out/tpc_sweep.jsonl; done
echo "=== tpc sweep: B16 Hkv1 ctx64K"
for t in 0 8 16; do echo "tpc=$t"; VATTN_DECODE_TPC=$t $B decode --hq 8 --hkv 1 --batch 16 --ctx 65536 | tee -a gpurun_out/tpc_sweep.jsonl; done
echo "=== headline shape sanity"; $B decode --ctx 32768
echo "=== done"
