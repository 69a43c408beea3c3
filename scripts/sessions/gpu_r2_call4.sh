#!/bin/bash
# Round 2, fourth GPU session: per-group stream-K prefill (+ MODE 2 softmax), ncu on the decode schedules
# at the small-batch shape, quick parity subset.  One GPU.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call4.log) 2>&1
B="timeout 300 python scripts/bench_extra.py"
echo "=== parity subset"
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_baseline_configs.py tests/test_gpu_allocator.py tests/test_gpu_oproj.py -q --timeout 300 --tb=short -k "prefill or pod or masked or lse or chunked or alias or oproj or gemm or graph" 2>&1 | tail -15
VATTN_PREFILL_MODE=2 timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_baseline_configs.py -q --timeout 300 --tb=short -k "prefill or masked or lse or chunked" 2>&1 | tail -8
VATTN_DECODE_SCHED=streamk timeout 600 python -m pytest tests/test_gpu_attention.py -q --timeout 300 --tb=short -k "decode" 2>&1 | tail -4
echo "=== prefill: stream-K per group (mode 1), mode 2, grid"
for c in 2048 512 8192; do $B prefill --chunk $c; VATTN_PREFILL_MODE=2 $B prefill --chunk $c; VATTN_PREFILL_SCHED=grid $B prefill --chunk $c; done
echo "=== ncu: decode schedules at B16 x Hkv1 x 32K"
for sched in grid streamk; do
  VATTN_DECODE_SCHED=$sched timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_ -s 6 -c 2 -o gpurun_out/r2_decode_small_$sched -f python scripts/bench_extra.py decode --hq 8 --hkv 1 --batch 16 --ctx 32768 --calls 4 > /dev/null 2>&1
done
echo "=== ncu: prefill stream-K kernel, chunk 2048 deep in the context"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill_sk -s 60 -c 1 -o gpurun_out/r2_prefill_sk -f python scripts/bench_extra.py prefill --chunk 2048 --iters 1 > /dev/null 2>&1
VATTN_PREFILL_MODE=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill_sk -s 60 -c 1 -o gpurun_out/r2_prefill_sk_mode2 -f python scripts/bench_extra.py prefill --chunk 2048 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
echo "=== done"
