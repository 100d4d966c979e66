#!/usr/bin/env python
"""Experiment: do two independent B=256 planner passes overlap on two streams (two contexts)?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import goal_of, synth_candidates
from ditreeonlineplanner_b200 import Context, load_maze, load_metadata
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init
grid = load_maze("boxes").astype(np.float32); meta = load_metadata("carmaze")
dims = UNET_DIMS["large"]; sd = random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctxs = []
for _ in range(2):
    c = Context(0); c.set_map(grid)
    c.load_denoiser(sd, action_dim=2, horizon=64, cond_dim=7, emb_dim=400, map_size=20, down_dims=dims, max_batch=B)
    ctxs.append(c)
st, prev = synth_candidates(grid, B, 1)
st = torch.as_tensor(st).cuda(); prev = torch.as_tensor(prev).cuda()
goal = torch.as_tensor(goal_of(grid).astype(np.float32)).cuda()
noise = torch.randn((B, 64, 2), device="cuda")
def one(c):
    lm = c.local_map(st, 20, 0.2, bf16_signed=True)
    cond = c.build_cond_car(st, prev, goal, meta, 20.0)
    a = c.fm_sample(noise, cond, lm, 1, meta["Actions_mean"], meta["Actions_std"])
    return c.propagate_collide(st, a, goal_of(grid), S=8, want_traj=True)
s = [torch.cuda.Stream(), torch.cuda.Stream()]
for k in range(2):
    with torch.cuda.stream(s[k]):
        for _ in range(3): one(ctxs[k])
torch.cuda.synchronize()
N = 40
t0 = time.perf_counter()
for _ in range(N):
    one(ctxs[0])
torch.cuda.synchronize()
t1 = (time.perf_counter() - t0) / N * 1e3
t0 = time.perf_counter()
for i in range(N):
    with torch.cuda.stream(s[i & 1]):
        one(ctxs[i & 1])
torch.cuda.synchronize()
t2 = (time.perf_counter() - t0) / N * 1e3
print(f"B={B}: one stream {t1:.3f} ms / pass; two streams {t2:.3f} ms / pass ({t1/t2:.2f}x)")
