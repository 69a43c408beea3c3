#!/bin/bash
# Round 2, final one-GPU session: whole parity suite, smoke, bench (all legs), launch list, one ncu
# --set full capture per kernel family, compute-sanitizer logs.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_configs.jsonl
exec > >(tee gpurun_out/r2_final.log) 2>&1
B="timeout 200 python scripts/bench_extra.py"
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short > gpurun_out/r2_final_pytest.txt 2>&1; tail -5 gpurun_out/r2_final_pytest.txt
echo "=== smoke"; timeout 200 python __graft_entry__.py smoke
echo "=== bench (ours)"; timeout 600 python bench.py | tee gpurun_out/r2_bench_final.json | cut -c1-400
echo "=== bench (reference arm)"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/r2_bench_reference.json | cut -c1-300
echo "=== bench --impl fa_vattn"; timeout 300 python bench.py --impl fa_vattn --steps 3 --warmup 2 2>/dev/null | grep '^{' | tee gpurun_out/r2_bench_fa_vattn.json
echo "=== secondary workloads"
for c in 2048 512 8192; do $B prefill --chunk $c; done | tee gpurun_out/r2_extra_prefill.jsonl
$B pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 64 --decode-len 16384 --iters 10 | tee gpurun_out/r2_extra_pod.jsonl
$B pod | tee -a gpurun_out/r2_extra_pod.jsonl
for c in 32768 65536 131072; do $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; done | tee gpurun_out/r2_extra_decode.jsonl
$B decode --ctx 131072 --ragged | tee -a gpurun_out/r2_extra_decode.jsonl
$B alloc | tee gpurun_out/r2_extra_alloc.jsonl
echo "=== launch list of the bench command (shares, not absolutes)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
echo "=== ncu --set full, one capture per kernel family"
N="timeout 240 ncu --set full --clock-control none --import-source on"
$N -k regex:decode_tc -s 40 -c 1 -o gpurun_out/r2_decode_tc -f python scripts/bench_extra.py decode --ctx 32768 --calls 4 > /dev/null 2>&1
$N -k regex:prefill2_tc -s 60 -c 1 -o gpurun_out/r2_prefill2_split -f python scripts/bench_extra.py prefill --chunk 2048 --iters 1 > /dev/null 2>&1
$N -k regex:pod_dual -s 2 -c 1 -o gpurun_out/r2_pod_dual -f python scripts/bench_extra.py pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 64 --decode-len 16384 --iters 1 > /dev/null 2>&1
$N -k regex:"oproj_allreduce|rope_qk|cache_flat_vec|combine_kernel" -c 8 -o gpurun_out/r2_small_kernels -f python -m pytest tests/test_gpu_oproj.py tests/test_gpu_attention.py -q -x -k "graph_replay or rotary_append or cache_flat or decode_split_counts" > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
echo "=== compute-sanitizer memcheck (decode, prefill incl. split items, POD dual-role, o_proj)"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_attention.py tests/test_gpu_oproj.py -q --timeout 450 -x -k "test_decode_split_counts or test_lse_output or (test_prefill_matches_oracle and auto) or test_pod_fused_many or gemm_matches" 2>&1 | tail -4 | tee gpurun_out/r2_sanitizer_memcheck.log
echo "=== compute-sanitizer racecheck (same kernels, smallest cases)"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_attention.py -q --timeout 280 -x -k "test_lse_output or test_fully_masked" 2>&1 | tail -4 | tee gpurun_out/r2_sanitizer_racecheck.log
echo "=== done"
