#!/bin/bash
# ncu evidence for profiles/ (one GPU, never under a multi-rank launch):
#   launch list of the bench step + full captures of the decode and prefill kernels.
set -u
mkdir -p gpurun_out/prof
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --resident-layers 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/prof/launches.csv $B > gpurun_out/prof/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_tc_kernel -s 100 -c 2 -f \
    -o gpurun_out/prof/decode_tc $B > gpurun_out/prof/ncu_decode.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prefill2_tc_kernel -s 40 -c 1 -f \
    -o gpurun_out/prof/prefill2_tc python scripts/bench_extra.py prefill --chunk 2048 --iters 1 \
    > gpurun_out/prof/ncu_prefill.log 2>&1
ls -la gpurun_out/prof
