"""tcgen05 GEMM core vs torch (bf16 inputs, fp32 accumulate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ditreeonlineplanner_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 64), (256, 256, 128), (1000, 128, 192), (4096, 512, 1536),
                                   (16384, 2048, 1024), (77, 64, 448)])
def test_gemm_bf16(ctx, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    got = ctx.gemm_bf16(a, w)
    torch.cuda.synchronize()
    want = a.float() @ w.float().t()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, f"max err {err} (scale {scale})"
