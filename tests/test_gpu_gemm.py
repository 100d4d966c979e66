"""tcgen05 GEMM core vs torch (bf16 inputs, fp32 accumulate)."""
import numpy as np
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


@pytest.mark.parametrize("B,H,Cin,N,k,s,p,resid,relu", [
    (1, 5, 64, 64, 3, 1, 1, False, True),      # one sample, 25 rows of a 128-row tile
    (20, 5, 64, 64, 3, 1, 1, True, True),      # 5 samples per tile (125 rows), several tiles, CTA pairs, residual
    (8, 5, 64, 128, 3, 2, 1, False, True),     # stride 2 = TMA traversal stride, 5x5 -> 3x3
    (8, 5, 64, 128, 1, 2, 0, False, False),    # the 1x1 stride-2 downsample branch (no ReLU)
    (30, 3, 128, 128, 3, 1, 1, True, True),    # 9 rows per sample: 14 samples per tile
    (40, 3, 128, 256, 3, 2, 1, False, True),   # 3x3 -> 2x2, two N tiles
    (70, 2, 256, 256, 3, 1, 1, True, True),    # 4 rows per sample (divides 128)
    (6, 10, 64, 64, 3, 1, 1, False, True),     # 100 rows per sample: one sample per tile, 28 empty rows
    (300, 5, 64, 64, 3, 1, 1, True, True),     # planner batch size, 60 tiles
])
def test_conv2d_gn_fused(ctx, B, H, Cin, N, k, s, p, resid, relu):
    """The encoder's fused launch -- 2-D conv through TMA tap addressing (zero padding = out-of-bounds fill, conv stride =
    traversal stride), GroupNorm over (pixels x 16 channels) for ANY pixel count <= 128, residual, ReLU -- against
    torch's conv2d + group_norm in fp32 on the same bf16 inputs; every sample on its own (a tile boundary bug corrupts
    single samples)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(B * 7 + H)
    x = torch.randn(B, H, H, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(torch.bfloat16)
    ga, be = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.2
    OH = (H + 2 * p - k) // s + 1
    r = torch.randn(B, OH, OH, N, generator=g).to(torch.bfloat16) if resid else None
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=s, padding=p)
    y = F.group_norm(y, N // 16, ga, be, 1e-5)
    if r is not None:
        y = y + r.float().permute(0, 3, 1, 2)
    if relu:
        y = F.relu(y)
    want = y.permute(0, 2, 3, 1).numpy()
    got = ctx.conv2d_gn(x.cuda(), w.reshape(N, -1).contiguous().cuda(), ga, be, k, s, p, None if r is None else r.cuda(),
                        relu).float().cpu().numpy()
    assert got.shape == want.shape
    worst = max(float(np.linalg.norm(got[b] - want[b]) / np.linalg.norm(want[b])) for b in range(B))
    assert worst < 6e-3, worst     # bf16 output rounding: 2^-9 relative per element
