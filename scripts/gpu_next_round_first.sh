#!/bin/bash
# Queued measurements that the round-1 GPU budget did not cover.  One GPU:
#   bash scripts/gpu_next_round_first.sh            (~4 min)
# then, on an 8-GPU box:  bash scripts/gpu_next_round_first.sh n8   (~3 min x 8)
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/next_round_first.log) 2>&1
B="timeout 200 python scripts/bench_extra.py"
if [ "${1:-}" = "n8" ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
  echo "=== fused GEMM + all-reduce parity / latency at world 8"
  timeout 300 $TR --master-port 29631 scripts/debug/tp_fused_test.py 2>&1 | grep -v "^\*\|OMP\|^$"
  echo "=== bench default (fused + graph)"; timeout 300 $TR --master-port 29632 bench.py --gpus 8 2>&1 | grep '^{' | tee gpurun_out/bench_tp8_fused_graph.json
  echo "=== bench nccl + graph"; timeout 300 $TR --master-port 29633 bench.py --gpus 8 --no-e2e --tp-collective nccl 2>&1 | grep '^{' | tee gpurun_out/bench_tp8_nccl_graph.json
  exit 0
fi
echo "=== prefill with the S row held in registers (setmaxnreg 56 / 224): parity, then throughput vs the default"
VATTN_PREFILL_REGS=1 timeout 300 python -m pytest tests/test_gpu_attention.py -q --timeout 60 -k "prefill or pod or masked or lse" 2>&1 | tail -4
for c in 512 2048 8192; do
  timeout 300 python scripts/bench_extra.py prefill --chunk $c
  VATTN_PREFILL_REGS=1 timeout 300 python scripts/bench_extra.py prefill --chunk $c
  VATTN_PREFILL_REGS=2 timeout 300 python scripts/bench_extra.py prefill --chunk $c
done
VATTN_PREFILL_REGS=2 timeout 300 python -m pytest tests/test_gpu_attention.py -q --timeout 60 -k "prefill or pod or masked or lse" 2>&1 | tail -4
echo "=== host-buffer entry points: parity (incl. the opt-in pipelined variant), then e2e with it"
VATTN_TEST_PIPELINED=1 timeout 200 python -m pytest tests/test_zz_gpu_host_path.py -q --timeout 60 2>&1 | tail -5
VATTN_E2E_PIPELINED=1 timeout 400 python bench.py --no-cpu | tee gpurun_out/bench_e2e_pipelined.json
echo "=== headline decode: tiles per chunk beyond the cap of 16 (model: 32..64 may save a few % of waves)"
for t in 0 16 32 64 128; do echo "tpc=$t"; VATTN_DECODE_TPC=$t $B decode --ctx 32768; done
echo "=== fused o_proj kernel, world 1, dense ring"; timeout 120 python scripts/debug/oproj_latency.py
echo "=== POD arms after the strategy switch (+ the co-resident 'lean' strategy: parity first)"
VATTN_TEST_POD_LEAN=1 timeout 200 python -m pytest tests/test_gpu_attention.py -q --timeout 60 -k "pod_fused_many" 2>&1 | tail -4
$B pod --prefills 1 --prefill-len 16384 --prefill-chunk 2048 --decodes 64 --decode-len 16384 --iters 10 --lean
$B pod --lean
echo "=== done"
