#!/usr/bin/env python
"""One planner chunk (local map + cond + K=1 sampler + 8-step propagate) at a given batch: for ncu launch lists."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import goal_of, synth_candidates
from ditreeonlineplanner_b200 import get_context, load_maze, load_metadata
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = get_context(0)
grid = load_maze("boxes").astype(np.float32); ctx.set_map(grid); meta = load_metadata("carmaze")
dims = UNET_DIMS["large"]
ctx.load_denoiser(random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims), action_dim=2, horizon=64,
                  cond_dim=7, emb_dim=400, map_size=20, down_dims=dims, max_batch=4096)
st, prev = synth_candidates(grid, B, 1)
st = torch.as_tensor(st).cuda(); prev = torch.as_tensor(prev).cuda()
goal = torch.as_tensor(goal_of(grid).astype(np.float32)).cuda()
noise = torch.randn((B, 64, 2), device="cuda")
for _ in range(reps):
    lm = ctx.local_map(st, 20, 0.2, bf16_signed=True)
    cond = ctx.build_cond_car(st, prev, goal, meta, 20.0)
    a = ctx.fm_sample(noise, cond, lm, 1, meta["Actions_mean"], meta["Actions_std"])
    r = ctx.propagate_collide(st, a, goal_of(grid), S=8, want_traj=True)
torch.cuda.synchronize()
print("ok", ctx.launches)
