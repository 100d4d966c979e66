#!/usr/bin/env python
"""Propagate+collide kernel alone: layouts x batch sizes, CUDA-event timed."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import goal_of, synth_candidates  # noqa: E402
from ditreeonlineplanner_b200 import Context, load_maze  # noqa: E402

grid = load_maze("boxes").astype(np.float32)
ctx = Context(0)
ctx.set_map(grid)
goal = goal_of(grid)
S = 50
ONLY = sys.argv[1] if len(sys.argv) > 1 else None   # e.g. "soa,no traj" : one config at B = 2^20 (for ncu)
BS = [int(x) for x in os.environ["PROP_B"].split(",")] if os.environ.get("PROP_B") else None
for B in (BS or ((1 << 20,) if ONLY else (4096, 1 << 16, 1 << 20, 1 << 22))):
    st_np, _ = synth_candidates(grid, B, 5)
    st = torch.as_tensor(st_np).cuda()
    act = torch.randn((B, S, 2), device="cuda") * torch.tensor([1.006, 0.923], device="cuda") + torch.tensor([0.451, 0.0], device="cuda")
    st_soa = st.t().contiguous()
    act_soa = act.permute(1, 2, 0).contiguous()
    for name, fn in (("rows+traj", lambda: ctx.propagate_collide(st, act, goal)),
                     ("rows,no traj", lambda: ctx.propagate_collide(st, act, goal, want_traj=False)),
                     ("soa+traj", lambda: ctx.propagate_collide(st_soa, act_soa, goal, soa=True)),
                     ("soa,no traj", lambda: ctx.propagate_collide(st_soa, act_soa, goal, soa=True, want_traj=False))):
        if ONLY and name != ONLY:
            continue
        for _ in range(1 if ONLY else 3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 1 if ONLY else (10 if B >= (1 << 20) else 50)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_edge = 4 * (2 * 6 + S * 2 + (S * 6 if "no traj" not in name else 0)) + 8
        print(f"B={B:8d} {name:13s} {ms*1e3:10.1f} us  {B/ms/1e3:9.1f} M edges/s  {B*bytes_edge/ms/1e6:8.1f} GB/s "
              f"({B*bytes_edge/ms/1e6/6547.2*100:5.1f}% of measured HBM peak)")
