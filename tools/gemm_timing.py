#!/usr/bin/env python
"""Cycle accounting of the GEMM kernel's warp roles for one layer shape (variant build with -DGEMM_TIMING):

    DITREE_LIB=.../libditree_timing.so DITREE_GEMM_DBG=512,1536,0 python tools/gemm_timing.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import goal_of, synth_candidates  # noqa: E402
from ditreeonlineplanner_b200 import Context, load_maze, load_metadata  # noqa: E402
from ditreeonlineplanner_b200.expansion import TreeExpander  # noqa: E402
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init  # noqa: E402

B = 4096
grid = load_maze("boxes").astype(np.float32)
ctx = Context(0)
ctx.set_map(grid)
dims = UNET_DIMS["large"]
ctx.load_denoiser(random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims), 2, 64, 7, 400, 20, dims, B)
exp = TreeExpander(ctx, load_metadata("carmaze"), 20, 0.2, num_diffusion_iters=2, action_horizon=50)
st, prev = synth_candidates(grid, B, 1000)
st, prev = torch.as_tensor(st).cuda(), torch.as_tensor(prev).cuda()
noise = torch.randn((B, 64, 2), device="cuda")
fn = ctx.lib.dt_gemm_timing
fn.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_uint64 * 16)()
exp.expand_device(st, prev, goal_of(grid), noise=noise)
torch.cuda.synchronize()
fn(None, 1)
exp.expand_device(st, prev, goal_of(grid), noise=noise)
torch.cuda.synchronize()
fn(buf, 0)
v = [int(x) for x in buf]
print("shape filter", os.environ.get("DITREE_GEMM_DBG"))
if v[0]:
    print(f"MMA warp: tiles {v[0]}, per tile: wait for a free accumulator {v[1] / v[0]:.0f} cyc, main loop {v[2] / v[0]:.0f} cyc")
if v[3]:
    print(f"epilogue warp 0: tiles {v[3]}, per tile: staging+barriers {v[8] / v[3]:.0f}, wait for accumulator {v[4] / v[3]:.0f}, "
          f"pass 1 {v[5] / v[3]:.0f}, stats+coefficients {v[6] / v[3]:.0f}, pass 2 {v[7] / v[3]:.0f} cyc")
if v[10]:
    print(f"kernel: {v[9] / v[10]:.0f} cycles per CTA over {v[10]} CTAs")
