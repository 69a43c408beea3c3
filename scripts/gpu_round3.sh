#!/bin/bash
# Third GPU pass: full parity suite, secondary workloads, ncu evidence, headline bench.
set -u
mkdir -p gpurun_out/prof
exec > >(tee gpurun_out/round3.log) 2>&1
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -s -k "not trace" 2>&1 | grep -vE "^\s*$" | tail -40
echo "=== prefill ours"; for c in 512 2048 8192; do timeout 600 python scripts/bench_extra.py prefill --chunk $c --impl ours; done
echo "=== prefill fa / fi"; timeout 600 python scripts/bench_extra.py prefill --chunk 2048 --impl fa; timeout 600 python scripts/bench_extra.py prefill --chunk 2048 --impl fi
echo "=== pod"; timeout 600 python scripts/bench_extra.py pod --impl ours; timeout 600 python scripts/bench_extra.py pod --impl fa
echo "=== alloc"; timeout 600 python scripts/bench_extra.py alloc
echo "=== bench ours"; timeout 900 python bench.py | tee gpurun_out/bench_ours3.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/prof/launches_r1.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --resident-layers 1 > gpurun_out/prof/launches_bench.log 2>&1
tail -2 gpurun_out/prof/launches_bench.log | cut -c1-300
echo "=== ncu full decode_tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_tc_kernel -s 100 -c 2 -f -o gpurun_out/prof/decode_tc_r1 \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --resident-layers 1 > gpurun_out/prof/ncu_decode.log 2>&1
tail -3 gpurun_out/prof/ncu_decode.log | cut -c1-300
echo "=== ncu full prefill_tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prefill_tc_kernel -s 40 -c 1 -f -o gpurun_out/prof/prefill_tc_r1 \
   python scripts/bench_extra.py prefill --chunk 2048 --iters 1 > gpurun_out/prof/ncu_prefill.log 2>&1
tail -3 gpurun_out/prof/ncu_prefill.log | cut -c1-300
ls -la gpurun_out/prof
echo "=== done"
