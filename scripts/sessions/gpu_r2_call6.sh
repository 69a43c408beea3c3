#!/bin/bash
# Round 2, sixth GPU session: split-item grid prefill, one-wave decode chunks + in-kernel combine for
# small problems, dual-role POD side timings.  One GPU.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call6.log) 2>&1
B="timeout 300 python scripts/bench_extra.py"
echo "=== parity (whole suite)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --tb=short > gpurun_out/r2_call6_pytest.txt 2>&1; tail -8 gpurun_out/r2_call6_pytest.txt; grep -n "^E  " gpurun_out/r2_call6_pytest.txt | head -20
echo "=== prefill: split grid (auto) vs no split"
for c in 2048 512 8192; do $B prefill --chunk $c; VATTN_PREFILL_SPLITS=1 $B prefill --chunk $c; done
VATTN_PREFILL_SPLITS=8 $B prefill --chunk 2048; VATTN_PREFILL_SPLITS=8 $B prefill --chunk 512
echo "=== decode small: one-wave chunks + in-kernel combine (default) vs separate combine"
for c in 32768 65536 131072; do $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; VATTN_DECODE_COMBINE_INKERNEL=0 $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; done
$B decode --hq 4 --hkv 1 --batch 64 --ctx 32768; VATTN_DECODE_COMBINE_INKERNEL=0 $B decode --hq 4 --hkv 1 --batch 64 --ctx 32768
$B decode --ctx 32768
echo "=== POD: hybrid; dual-role kernel with only one side loaded"
$B pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 64 --decode-len 16384 --iters 10
$B pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 1 --decode-len 256 --iters 10
$B pod --prefills 1 --prefill-len 256 --prefill-chunk 256 --decodes 64 --decode-len 16384 --iters 10
$B pod
echo "=== done"
