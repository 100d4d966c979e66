#!/usr/bin/env python
"""Flat split-K (encoder-tail GEMM shapes) on/off: time and max error vs fp32 matmul."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ditreeonlineplanner_b200 import Context
ctx = Context(0)
torch.manual_seed(0)
for (M, N, K) in ((256, 512, 4608), (256, 512, 2304), (1024, 256, 2304), (1024, 256, 1152), (2304, 128, 1152), (300, 512, 4608), (64, 512, 4608), (100, 128, 1152)):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    ref = a.float() @ w.float().t()
    line = f"M={M:5d} N={N:4d} K={K:5d}:"
    for on in (0, 1):
        ctx.set_option("splitk", on)
        out = ctx.gemm_bf16(a, w)
        err = ((out - ref).abs().max() / ref.abs().max()).item()
        for _ in range(5): ctx.gemm_bf16(a, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): ctx.gemm_bf16(a, w)
        e1.record(); torch.cuda.synchronize()
        line += f"  splitk={on}: {e0.elapsed_time(e1) / 50 * 1e3:6.1f} us err {err:.1e}"
    print(line)
