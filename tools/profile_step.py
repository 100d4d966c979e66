#!/usr/bin/env python
"""One warm-up expansion + one measured expansion of the bench workload, for ncu."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import goal_of, synth_candidates  # noqa: E402
from ditreeonlineplanner_b200 import Context, load_maze, load_metadata  # noqa: E402
from ditreeonlineplanner_b200.expansion import TreeExpander  # noqa: E402
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--ode-steps", type=int, default=10)
ap.add_argument("--rollout", type=int, default=50)
ap.add_argument("--denoiser", default="large")
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--csv", default=None, help="write per-GEMM-launch timings (CUDA events) to this CSV")
a = ap.parse_args()
grid = load_maze("boxes").astype(np.float32)
ctx = Context(0)
ctx.set_map(grid)
dims = UNET_DIMS[a.denoiser]
ctx.load_denoiser(random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims), 2, 64, 7, 400, 20, dims, a.batch)
exp = TreeExpander(ctx, load_metadata("carmaze"), 20, 0.2, num_diffusion_iters=a.ode_steps, action_horizon=a.rollout)
st, prev = synth_candidates(grid, a.batch, 1000)
st, prev = torch.as_tensor(st).cuda(), torch.as_tensor(prev).cuda()
noise = torch.randn((a.batch, 64, 2), device="cuda")
for i in range(a.passes):
    if i == a.passes - 1 and a.csv:
        ctx.profile_begin()
    res = exp.expand_device(st, prev, goal_of(grid), noise=noise)
torch.cuda.synchronize()
if a.csv:
    ms, n = ctx.profile_end()
    ctx.profile_csv(a.csv)
    print("gemm ms", ms, "launches", n)
print("launches", ctx.launches, "ok edges", int((res["first_coll"] < 0).sum()))
