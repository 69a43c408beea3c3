"""GPU: the fused o_proj GEMM + all-reduce kernel (csrc/oproj_allreduce.cu) with world = 1 -- the
tcgen05 GEMM, the tile push / flag / reduce epilogue against the rank's own slot, epoch handling
across calls and under CUDA-graph replay.  Reference: fp32 matmul of the same 16-bit operands,
rounded once.  Tolerance: the kernel accumulates in fp32 on the tensor core (order differs from
torch's), so |err| <= 2 ulp of the output dtype relative to the output scale.  The multi-rank
exchange is covered by tests/test_gpu_tp_peer.py (needs 2 GPUs)."""
import pytest
import torch

from vattention_b200.tp import FusedOProjAllReduce

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def ref_gemm(x, w):
    return (x.float() @ w.float().t()).to(x.dtype)


def check(got, want, dtype):
    ulp = {torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10}[dtype]
    scale = want.float().abs().max().item()
    err = (got.float() - want.float()).abs().max().item()
    assert err <= 2 * ulp * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.timeout(120)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("tokens,hidden,k", [(64, 4096, 512), (1, 4096, 2048), (50, 1024, 128), (128, 4096, 4096),
                                             (17, 256, 64), (96, 8192, 1024), (64, 11008, 256), (8, 160, 64)])
def test_gemm_matches_fp32_reference(dtype, tokens, hidden, k):
    g = torch.Generator(device=DEV).manual_seed(tokens * 7 + k)
    w = (torch.randn(hidden, k, device=DEV, generator=g) * 0.05).to(dtype)
    op = FusedOProjAllReduce(w, 128, local_only=True)
    for call in range(3):                                  # epochs 1..3: both parities, slot reuse
        x = torch.randn(tokens, k, device=DEV, generator=g).to(dtype)
        got = op(x).clone()
        torch.cuda.synchronize()
        check(got, ref_gemm(x, w), dtype)
    assert not op.failed()
    assert op.epoch_state[4].item() == 3          # per-tile epoch (tile 0): one per call


@pytest.mark.timeout(120)
def test_strided_input_and_smaller_batches_share_buffers():
    g = torch.Generator(device=DEV).manual_seed(1)
    w = (torch.randn(2048, 1024, device=DEV, generator=g) * 0.05).bfloat16()
    op = FusedOProjAllReduce(w, 64, local_only=True)
    big = torch.randn(64, 4096, device=DEV, generator=g).bfloat16()
    for tokens in (64, 3, 33, 64):
        x = big[:tokens, 1024:2048]                        # row stride 4096, 16-byte aligned offset
        got = op(x).clone()
        check(got, ref_gemm(x, w), torch.bfloat16)
    assert not op.failed()


@pytest.mark.timeout(120)
def test_cuda_graph_replay_advances_the_epoch_on_device():
    g = torch.Generator(device=DEV).manual_seed(2)
    w = (torch.randn(4096, 512, device=DEV, generator=g) * 0.05).bfloat16()
    op = FusedOProjAllReduce(w, 64, local_only=True)
    x = torch.randn(64, 512, device=DEV, generator=g).bfloat16()
    op(x)                                                  # warm-up outside the graph (module load)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y1 = op(x).clone()
        y2 = op(x).clone()                                 # two calls per replay: both parities
    for rep in range(3):
        x.copy_(torch.randn(64, 512, device=DEV, generator=g).bfloat16())
        graph.replay()
        torch.cuda.synchronize()
        want = ref_gemm(x, w)
        check(y1, want, torch.bfloat16)
        assert torch.equal(y1, y2)
    assert op.epoch_state[4].item() == 1 + 2 * 3 and not op.failed()


def test_argument_rules():
    w = torch.zeros(4096, 512, device=DEV, dtype=torch.bfloat16)
    op = FusedOProjAllReduce(w, 128, local_only=True)
    with pytest.raises(ValueError):
        op(torch.zeros(4, 256, device=DEV, dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        FusedOProjAllReduce(w, 256, local_only=True)
    with pytest.raises(RuntimeError, match="multiple of 32"):
        FusedOProjAllReduce(torch.zeros(200, 512, device=DEV, dtype=torch.bfloat16), 64, local_only=True)(
            torch.zeros(4, 512, device=DEV, dtype=torch.bfloat16))
