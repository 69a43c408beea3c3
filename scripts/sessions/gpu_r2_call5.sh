#!/bin/bash
# Round 2, fifth GPU session: ping-pong token in the prefill softmax (grid / stream-K x MODE 1 / 2),
# dual-role POD kernel (parity, then arms).  One GPU.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call5.log) 2>&1
B="timeout 300 python scripts/bench_extra.py"
echo "=== parity subset (default), then MODE 2, then grid schedule"
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_baseline_configs.py -q --timeout 300 --tb=short -k "prefill or pod or masked or lse or chunked" 2>&1 | tail -12
VATTN_PREFILL_MODE=2 timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_baseline_configs.py -q --timeout 300 --tb=short -k "prefill or masked or lse or chunked" 2>&1 | tail -5
VATTN_PREFILL_SCHED=grid timeout 900 python -m pytest tests/test_gpu_attention.py -q --timeout 300 --tb=short -k "prefill or masked or lse" 2>&1 | tail -3
echo "=== prefill: [stream-K mode 1, stream-K mode 2, grid mode 1]"
for c in 2048 512 8192; do $B prefill --chunk $c; VATTN_PREFILL_MODE=2 $B prefill --chunk $c; VATTN_PREFILL_SCHED=grid $B prefill --chunk $c; done
echo "=== POD arms (serial / two streams / 15 / persistent 9 / dual-role 64)"
$B pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 64 --decode-len 16384 --iters 10
$B pod
echo "=== ncu: grid prefill kernel with the token, chunk 2048 deep in the context"
VATTN_PREFILL_SCHED=grid timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill2_tc -s 60 -c 1 -o gpurun_out/r2_prefill2_grid_token -f python scripts/bench_extra.py prefill --chunk 2048 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/r2_prefill2_grid_token.ncu-rep
echo "=== done"
