#!/bin/bash
# Round 2, eighth GPU session: prefill3 after the phase-aliasing fix -- watchdog parity first (aborts on
# failure), then throughput and ncu.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call8.log) 2>&1
B="timeout 120 python scripts/bench_extra.py"
echo "=== prefill3 parity under the mbarrier watchdog"
VATTN_B200_LIB=$PWD/vattention_b200/libvattn_b200_dbg.so VATTN_PREFILL_KERNEL=3 timeout 300 python -m pytest tests/test_gpu_attention.py -q -x --timeout 60 --tb=short -k "prefill or masked or lse or wide_pitch or rotary" > gpurun_out/r2_call8_dbg.txt 2>&1
tail -4 gpurun_out/r2_call8_dbg.txt; grep -c watchdog gpurun_out/r2_call8_dbg.txt
if ! grep -q " passed" gpurun_out/r2_call8_dbg.txt || grep -q "failed\|watchdog" gpurun_out/r2_call8_dbg.txt; then echo "ABORT: prefill3 not correct"; grep watchdog gpurun_out/r2_call8_dbg.txt | sed "s/block ([0-9,]*) thread [0-9]*/block T/" | sort | uniq -c | head; exit 0; fi
echo "=== prefill3 parity, release build, full-size configs"
VATTN_PREFILL_KERNEL=3 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_baseline_configs.py tests/test_gpu_allocator.py -q --timeout 120 --tb=short -k "prefill or pod or masked or lse or chunked or wide_pitch or megacache or rotary" 2>&1 | tail -6
echo "=== prefill: kernel 3 vs kernel 2"
for c in 2048 512 8192; do VATTN_PREFILL_KERNEL=3 $B prefill --chunk $c; $B prefill --chunk $c; done
VATTN_PREFILL_KERNEL=3 VATTN_PREFILL_SPLITS=1 $B prefill --chunk 2048
echo "=== ncu: prefill3"
VATTN_PREFILL_KERNEL=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:prefill3_tc -s 60 -c 1 -o gpurun_out/r2_prefill3 -f python scripts/bench_extra.py prefill --chunk 2048 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/r2_prefill3.ncu-rep
echo "=== done"
