#!/usr/bin/env python
"""A/B of a launch option -- programmatic dependent launch (dt_set_option "pdl", default) or EXP_OPT=fork -- on one planner pass (local map + cond + K = 1 sampler +
8-step propagate) at small batches: device time per pass, bit-identity of the actions, and the sampler call alone
through its CUDA-graph replay (B <= 64)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import goal_of, synth_candidates
from ditreeonlineplanner_b200 import get_context, load_maze, load_metadata
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init, denoiser_flops
ctx = get_context(0)
grid = load_maze("boxes").astype(np.float32); ctx.set_map(grid); meta = load_metadata("carmaze")
dims = UNET_DIMS["large"]
ctx.load_denoiser(random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims), action_dim=2, horizon=64,
                  cond_dim=7, emb_dim=400, map_size=20, down_dims=dims, max_batch=4096)
enc, unet = denoiser_flops(1, down_dims=dims)
batches = [int(b) for b in os.environ.get("SB_BATCHES", "1,16,64,256,1024,4096").split(",")]
reps = int(os.environ.get("SB_REPS", "30"))
ref = {}
OPT = os.environ.get("EXP_OPT", "pdl")   # "pdl" or "fork" (side stream for the residual 1 x 1 convs)
for pdl in (0, 1, 0, 1):
    ctx.set_option(OPT, pdl)
    for B in batches:
        st, prev = synth_candidates(grid, B, 1)
        st = torch.as_tensor(st).cuda(); prev = torch.as_tensor(prev).cuda()
        goal = torch.as_tensor(goal_of(grid).astype(np.float32)).cuda()
        noise = torch.randn((B, 64, 2), device="cuda", generator=torch.Generator("cuda").manual_seed(B))
        def one():
            lm = ctx.local_map(st, 20, 0.2, bf16_signed=True)
            cond = ctx.build_cond_car(st, prev, goal, meta, 20.0)
            a = ctx.fm_sample(noise, cond, lm, 1, meta["Actions_mean"], meta["Actions_std"])
            return a, ctx.propagate_collide(st, a, goal_of(grid), S=8, want_traj=True)
        for _ in range(5): a, r = one()
        torch.cuda.synchronize()
        a = a.clone()
        if B in ref:
            same = bool(torch.equal(a, ref[B]))
        else:
            ref[B] = a; same = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(reps): one()
        e1.record(); t_host = (time.perf_counter() - t0) / reps * 1e3
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{OPT}={pdl} B={B:5d}: device {ms:7.3f} ms / pass, host enqueue {t_host:6.3f} ms, "
              f"{B*(enc+unet)/ms/1e9:7.1f} TFLOP/s, actions identical to the first run: {same}", flush=True)
