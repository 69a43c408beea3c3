#!/bin/bash
# Round 2, short one-GPU session: the headline line with the per-N crossing window, a steady-state
# launch list of the same command (set-up fills filtered out by name) and the missing --set full
# capture of decode_tc_kernel at the headline shape.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_decode_capture.log) 2>&1
echo "=== bench (ours, N = 1, no extras)"
timeout 200 python bench.py --no-extras --no-cpu | tee gpurun_out/r2_bench_n1_window.json | cut -c1-600
echo "=== launch list of the bench command (steady state)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"decode_tc|combine|nvjet|gemm|gemv|cutlass|oproj|append|rope" -s 40 -c 300 --csv \
  --log-file gpurun_out/r2_launches_bench_steady.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
wc -l gpurun_out/r2_launches_bench_steady.csv
echo "=== ncu --set full: decode_tc_kernel at the headline shape"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_tc -s 6 -c 1 \
  -o gpurun_out/r2_decode_tc_headline -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras --no-graph > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
echo "=== done"
