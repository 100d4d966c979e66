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


@pytest.mark.parametrize("M,N,K", [(256, 512, 4608), (300, 512, 4608), (64, 512, 4608), (1000, 256, 4096), (8, 64, 8192)])
def test_gemm_split_k_shapes(ctx, M, N, K):
    """Few tiles and a long K: the split-K work items + the reduce kernel (one-tile and flat multi-tile
    forms), against fp32 matmul and against the unsplit kernel (`splitk` option off)."""
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    want = a.float() @ w.float().t()
    scale = want.abs().max().item()
    try:
        ctx.set_option("splitk", 1)
        split = ctx.gemm_bf16(a, w)
        ctx.set_option("splitk", 0)
        plain = ctx.gemm_bf16(a, w)
    finally:
        ctx.set_option("splitk", 1)
    torch.cuda.synchronize()
    assert (split - want).abs().max().item() <= 1e-4 * scale
    assert (plain - want).abs().max().item() <= 1e-4 * scale
    assert (split - plain).abs().max().item() <= 2e-5 * scale   # same products, another summation order
