#!/usr/bin/env python
"""Time the scenario suite on the device-resident multi-scenario planner (1 GPU): scenarios/s, host share."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ditreeonlineplanner_b200 import scenarios as sc  # noqa: E402
from ditreeonlineplanner_b200.policies.fm_policy import DiffusionSampler  # noqa: E402
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--runs", type=int, default=10)
ap.add_argument("--unit-slots", type=int, default=8)
ap.add_argument("--denoiser", default="large")
ap.add_argument("--cap", type=int, default=4096)
ap.add_argument("--engine", default="device")
ap.add_argument("--streams", type=int, default=1)
a = ap.parse_args()
dims = UNET_DIMS[a.denoiser]
sd = random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims)
smp = DiffusionSampler(sd, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2, obs_history=1,
                       action_history=1, goal_conditioned=True, num_diffusion_iters=1, local_map_size=20,
                       max_batch=max(256, a.unit_slots * 256)).eval()
kw = {"unit_slots": a.unit_slots, "iteration_cap": a.cap, "streams": a.streams} if a.engine == "device" else \
    {"batch_size": 256, "iteration_cap": a.cap}
# warm-up: one unit
sc.run_suite(smp, total_runs=1, time_budget=1e9, planner_kwargs=dict(kw, **({"iteration_cap": 512})), engine=a.engine) \
    if a.engine == "device" else sc.run_car_unit(sc.load_scenarios("test_scenarios_car")[0], 0, 0, smp, 1e9, kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
table, _ = sc.run_suite(smp, total_runs=a.runs, time_budget=1e9, planner_kwargs=kw, engine=a.engine)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
rows = np.array(list(table.values()))
out = {"engine": a.engine, "units": len(table), "seconds": dt, "scenarios_per_s": len(table) / dt,
       "success_rows": int((rows[:, 1] == 1).sum()), "mean_tree_nodes": float(np.mean(rows[:, 6][rows[:, 6] > 0])),
       "mean_iterations": float(rows[:, 7].mean()), "stats": dict(sc.LAST_SUITE_STATS)}
print(json.dumps(out))
