#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_dbg.log) 2>&1
export VATTN_B200_LIB=$PWD/vattention_b200/libvattn_b200_dbg.so
VATTN_PREFILL_KERNEL=3 timeout 120 python -m pytest tests/test_gpu_attention.py -q -x --timeout 100 --tb=short -k "test_prefill_matches_oracle and auto and 1-256-32" > gpurun_out/r2_dbg_full.txt 2>&1
grep "\[p3\]" gpurun_out/r2_dbg_full.txt | sort | uniq -c | head -80
grep "block (0,9,0)" gpurun_out/r2_dbg_full.txt | head
echo "=== done"
