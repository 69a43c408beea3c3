#!/bin/bash
# Round 2, first GPU session: the full parity suite (incl. the new full-size BASELINE configs and the
# real-driver prefix-sharing test), then the measurements round 1 left queued.  One GPU, ~12 min.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_configs.jsonl
exec > >(tee gpurun_out/r2_call1.log) 2>&1
B="timeout 300 python scripts/bench_extra.py"
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15
echo "=== prefill REGS variants: parity"
VATTN_PREFILL_REGS=1 timeout 300 python -m pytest tests/test_gpu_attention.py -q --timeout 60 -k "prefill or pod or masked or lse" 2>&1 | tail -3
VATTN_PREFILL_REGS=2 timeout 300 python -m pytest tests/test_gpu_attention.py -q --timeout 60 -k "prefill or pod or masked or lse" 2>&1 | tail -3
echo "=== prefill throughput: default / REGS=1 / REGS=2"
for c in 2048 512; do
  $B prefill --chunk $c
  VATTN_PREFILL_REGS=1 $B prefill --chunk $c
  VATTN_PREFILL_REGS=2 $B prefill --chunk $c
done
echo "=== host-buffer entry points (pipelined variant): parity, then e2e"
VATTN_TEST_PIPELINED=1 timeout 200 python -m pytest tests/test_zz_gpu_host_path.py -q --timeout 60 2>&1 | tail -3
VATTN_E2E_PIPELINED=1 timeout 400 python bench.py --no-cpu | tee gpurun_out/r2_bench_e2e_pipelined.json
timeout 400 python bench.py --no-cpu | tee gpurun_out/r2_bench_default.json
echo "=== decode tiles per chunk on the headline shape"
for t in 16 32 64; do echo "tpc=$t"; VATTN_DECODE_TPC=$t $B decode --ctx 32768; done
echo "=== small-batch decode (configs[4] per-GPU shapes), baseline for the stream-K work"
for c in 32768 65536 131072; do $B decode --hq 8 --hkv 1 --batch 16 --ctx $c; done
$B decode --hq 4 --hkv 1 --batch 64 --ctx 32768
echo "=== POD: lean parity, then arms"
VATTN_TEST_POD_LEAN=1 timeout 200 python -m pytest tests/test_gpu_attention.py -q --timeout 60 -k "pod_fused_many" 2>&1 | tail -3
$B pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 64 --decode-len 16384 --iters 10 --lean
$B pod --lean
echo "=== done"
