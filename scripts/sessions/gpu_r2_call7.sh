#!/bin/bash
# Round 2, seventh GPU session: prefill3 (64-key tiles, double-buffered S): parity, throughput, ncu.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call7.log) 2>&1
B="timeout 300 python scripts/bench_extra.py"
echo "=== parity, VATTN_PREFILL_KERNEL=3"
VATTN_PREFILL_KERNEL=3 timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_baseline_configs.py tests/test_gpu_allocator.py -q --timeout 300 --tb=short -k "prefill or pod or masked or lse or chunked or wide_pitch or megacache or rotary" 2>&1 | tail -15
echo "=== prefill: kernel 3 vs kernel 2"
for c in 2048 512 8192; do VATTN_PREFILL_KERNEL=3 $B prefill --chunk $c; $B prefill --chunk $c; done
VATTN_PREFILL_KERNEL=3 VATTN_PREFILL_SPLITS=1 $B prefill --chunk 2048
VATTN_PREFILL_KERNEL=3 $B pod --prefills 8 --decodes 1 --decode-len 256 --iters 2
echo "=== ncu: prefill3"
VATTN_PREFILL_KERNEL=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill3_tc -s 60 -c 1 -o gpurun_out/r2_prefill3 -f python scripts/bench_extra.py prefill --chunk 2048 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/r2_prefill3.ncu-rep
echo "=== done"
