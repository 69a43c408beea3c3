#!/bin/bash
# Round 2, third GPU session: stream-K prefill (first run), decode stream-K with the parallel reducer,
# bench.py (graph at N = 1), failing tests with full tracebacks.  One GPU.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_configs.jsonl
exec > >(tee gpurun_out/r2_call3.log) 2>&1
B="timeout 300 python scripts/bench_extra.py"
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --tb=short > gpurun_out/r2_call3_pytest.txt 2>&1; tail -12 gpurun_out/r2_call3_pytest.txt
grep -n "Error\|error\|assert" gpurun_out/r2_call3_pytest.txt | head -40
echo "=== prefill: stream-K (default) vs grid"
for c in 2048 512 8192; do $B prefill --chunk $c; VATTN_PREFILL_SCHED=grid $B prefill --chunk $c; done
echo "=== decode: stream-K (default) vs grid"
for c in 32768 131072; do $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; VATTN_DECODE_SCHED=grid $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; done
$B decode --hq 4 --hkv 1 --batch 64 --ctx 32768; VATTN_DECODE_SCHED=grid $B decode --hq 4 --hkv 1 --batch 64 --ctx 32768
$B decode --ctx 32768; VATTN_DECODE_SCHED=grid $B decode --ctx 32768
echo "=== bench (new)"; timeout 900 python bench.py | tee gpurun_out/r2_bench_call3.json
echo "=== compute-sanitizer memcheck: decode stream-K + prefill stream-K (small shapes)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_attention.py -q --timeout 800 -x -k "test_decode_split_counts or test_lse_output or test_config2_shape_properties or (test_prefill_matches_oracle and auto)" 2>&1 | tail -8
echo "=== done"
