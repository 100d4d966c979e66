#!/usr/bin/env python
"""Experiment: where does the trajectory-store cost of propagate+collide come from?
Calls the C ABI directly with custom trajectory strides (SoA layout)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import goal_of, synth_candidates
from ditreeonlineplanner_b200 import Context, load_maze
from ditreeonlineplanner_b200 import _lib as L
from ditreeonlineplanner_b200.runtime import _ptr

grid = load_maze("boxes").astype(np.float32)
ctx = Context(0); ctx.set_map(grid); goal = goal_of(grid)
B, S = 1 << 20, 50
st_np, _ = synth_candidates(grid, B, 5)
st = torch.as_tensor(st_np).cuda().t().contiguous()
act = (torch.randn((B, S, 2), device="cuda") * torch.tensor([1.006, 0.923], device="cuda") + torch.tensor([0.451, 0.0], device="cuda")).permute(1, 2, 0).contiguous()
final = torch.empty_like(st); first = torch.empty(B, dtype=torch.int32, device="cuda"); done = torch.empty_like(first)
traj = torch.empty((S, 6, B), device="cuda")

def run(tptr, t_step, t_comp, label):
    def fn():
        rc = ctx.lib.dt_propagate_collide(ctx.h, _ptr(st), 1, B, _ptr(act), 1, 2 * B, B, B, S, float(goal[0]), float(goal[1]),
                                          tptr, 1, t_step, t_comp, _ptr(final), _ptr(first), _ptr(done), 1, ctx._stream())
        assert rc == 0
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:50s} {e0.elapsed_time(e1)/10*1e3:8.1f} us")

run(None, 6 * B, B, "no trajectory")
run(_ptr(traj), 6 * B, B, "full trajectory (S,6,B): 1.26 GB")
run(_ptr(traj), 0, B, "every step overwrites the same (6,B) block: 25 MB")
run(_ptr(traj), 0, 0, "every store of a lane hits the same (B,) row: 4 MB")
