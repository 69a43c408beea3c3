"""GPU (needs >= 2 devices, otherwise skipped): the peer-memory all-reduce kernel against an fp32 sum
of the gathered partials, launched the way the driver launches multi-GPU jobs (torchrun, NCCL)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_peer_allreduce_matches_fp32_sum_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577",
                        str(ROOT / "scripts" / "debug" / "tp_peer_test.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bitwise over 50 calls" in r.stdout


def test_fused_oproj_allreduce_world2():
    """GEMM + all-reduce in one kernel: parity with the sum of the ranks' bf16 partials and identical
    bits on every rank (scripts/debug/tp_fused_test.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29578",
                        str(ROOT / "scripts" / "debug" / "tp_fused_test.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "parity ok over 40 calls" in r.stdout
