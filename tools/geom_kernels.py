#!/usr/bin/env python
"""Geometry / reduction kernels at large batch, CUDA-event timed (and the target of the ncu captures under
profiles/): collision (13 B / state), local maps, lidar, nearest neighbour, probability-map draws."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ditreeonlineplanner_b200 import Context, load_maze

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ctx = Context(0)
grid = load_maze("boxes").astype(np.float32); ctx.set_map(grid)
rng = np.random.default_rng(0)
out = {}
N = 1 << 24
st = torch.as_tensor(np.stack([rng.uniform(-10, 10, N), rng.uniform(-10, 10, N), rng.uniform(-3.2, 3.2, N)], 1).astype(np.float32)).cuda()
soa = st.t().contiguous()
ms = timed(lambda: ctx.collide_car(st))
out["collide_car_rows"] = {"states": N, "ms": ms, "states_per_s": N / ms * 1e3, "GBps_13B": N * 13 / ms / 1e6}
P = 1 << 18
ms = timed(lambda: ctx.local_map(st[:P], 20, 0.2, bf16_signed=True))
out["local_map_bf16"] = {"poses": P, "ms": ms, "maps_per_s": P / ms * 1e3, "GBps": P * (12 + 800) / ms / 1e6}
ms = timed(lambda: ctx.local_map(st[:P], 20, 0.2))
out["local_map_f32"] = {"poses": P, "ms": ms, "maps_per_s": P / ms * 1e3, "GBps": P * (12 + 1600) / ms / 1e6}
L = 1 << 15
poses = torch.as_tensor(np.stack([rng.uniform(1, 19, L), rng.uniform(1, 19, L), rng.uniform(-3, 3, L)], 1).astype(np.float32)).cuda()
ms = timed(lambda: ctx.lidar_scan(poses), reps=10)
out["lidar"] = {"poses": L, "rays": L * 181, "ms": ms, "rays_per_s": L * 181 / ms * 1e3}
n, Q = 100_000, 1 << 16
nx = torch.rand(n, device="cuda") * 20 - 10; ny = torch.rand(n, device="cuda") * 20 - 10
q = torch.rand((Q, 2), device="cuda") * 20 - 10
ms = timed(lambda: ctx.nearest(nx, ny, q), reps=5)
out["nearest"] = {"nodes": n, "queries": Q, "ms": ms, "pairs_per_s": n * Q / ms * 1e3}
ms = timed(lambda: ctx.nearest_k(nx, ny, q[:8192], 8), reps=5)
out["nearest_k8"] = {"nodes": n, "queries": 8192, "ms": ms, "pairs_per_s": n * 8192 / ms * 1e3}
prob = ctx.edt_prior()
u = torch.rand(1 << 22, device="cuda", dtype=torch.float64)
ms = timed(lambda: ctx.sample_cells(prob, u))
out["sample_cells"] = {"draws": 1 << 22, "ms": ms, "draws_per_s": (1 << 22) / ms * 1e3}
print(json.dumps(out, indent=1))
