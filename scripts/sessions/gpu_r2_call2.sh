#!/bin/bash
# Round 2, second GPU session: stream-K decode + new bench.py + the parity suite without -x, pipe
# throughput micro-benchmark, sanitizer on the new kernel.  One GPU, ~13 min.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_configs.jsonl
exec > >(tee gpurun_out/r2_call2.log) 2>&1
B="timeout 300 python scripts/bench_extra.py"
echo "=== pipe throughput"; timeout 120 scripts/debug/pipe_throughput
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40
echo "=== small-batch decode, stream-K (default) vs grid"
for c in 32768 65536 131072; do $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; VATTN_DECODE_SCHED=grid $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; done
$B decode --hq 4 --hkv 1 --batch 64 --ctx 32768; VATTN_DECODE_SCHED=grid $B decode --hq 4 --hkv 1 --batch 64 --ctx 32768
$B decode --ctx 32768; VATTN_DECODE_SCHED=grid $B decode --ctx 32768
$B decode --ctx 131072 --ragged; VATTN_DECODE_SCHED=grid $B decode --ctx 131072 --ragged
echo "=== bench (new)"; timeout 900 python bench.py | tee gpurun_out/r2_bench_call2.json
echo "=== bench --impl fa_vattn (reference extension + flash_attn)"; timeout 300 python bench.py --impl fa_vattn --steps 3 --warmup 2 | tee gpurun_out/r2_bench_fa_vattn.json
echo "=== compute-sanitizer memcheck: decode stream-K + prefill (small shapes)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_attention.py -q --timeout 600 -x -k "decode_append_matches_oracle and auto and bfloat16 or decode_split_counts and auto and float16 or prefill_matches_oracle and auto and bfloat16" 2>&1 | tail -15
echo "=== done"
