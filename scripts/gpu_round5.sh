#!/bin/bash
# Round-1 late pass: rotary goldens + new parity cases, then the SURVEY 8(d) variants not yet measured
# (ragged / fp16 decode, TP-8 per-GPU shapes, 256 KB logical pages, two-stream POD arm).
set -u
mkdir -p gpurun_out/golden
exec > >(tee gpurun_out/round5.log) 2>&1
echo "=== goldens"; timeout 300 python oracle/gen_attn_golden.py gpurun_out/golden
echo "=== pytest (new cases)"
timeout 600 python -m pytest tests/test_gpu_attention.py -q --timeout 120 -k "rotary" 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_allocator.py -q --timeout 120 -k "decode_attention_over_virtual" 2>&1 | tail -5
B="timeout 300 python scripts/bench_extra.py"
echo "=== decode variants"
$B decode --ragged                          | tee -a gpurun_out/extra_v3.jsonl
$B decode --ragged --impl fa                | tee -a gpurun_out/extra_v3.jsonl
$B decode --dtype fp16                      | tee -a gpurun_out/extra_v3.jsonl
for ctx in 32768 65536 131072; do
  $B decode --hq 8 --hkv 1 --batch 16 --ctx $ctx          | tee -a gpurun_out/extra_v3.jsonl
done
$B decode --hq 8 --hkv 1 --batch 16 --ctx 131072 --impl fa | tee -a gpurun_out/extra_v3.jsonl
echo "=== prefill, 256 KB logical pages"
$B prefill --chunk 2048 --page-kb 256       | tee -a gpurun_out/extra_v3.jsonl
echo "=== pod arms"
$B pod                                      | tee -a gpurun_out/extra_v3.jsonl
$B pod --impl fa                            | tee -a gpurun_out/extra_v3.jsonl
$B pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 64 --decode-len 16384 --iters 10 | tee -a gpurun_out/extra_v3.jsonl
echo "=== done"
