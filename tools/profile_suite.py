#!/usr/bin/env python
"""Where does a scenario-suite unit spend its time?  cProfile of a few (scenario, run) units of the batched
planner + device-busy time from the library's own kernel-event profiler."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ditreeonlineplanner_b200 import get_context, scenarios as sc
from ditreeonlineplanner_b200.policies.fm_policy import DiffusionSampler
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init

if os.environ.get("OLD_RRT"):  # A/B: load a saved copy of the planner module in place of the package's
    import importlib.util
    import ditreeonlineplanner_b200.planners  # noqa: F401
    spec = importlib.util.spec_from_file_location("ditreeonlineplanner_b200.planners.RRT", os.environ["OLD_RRT"])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ditreeonlineplanner_b200.planners.RRT"] = mod
    spec.loader.exec_module(mod)
ctx = get_context(0)
sd = random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=UNET_DIMS["large"])
sampler = DiffusionSampler(sd, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2, obs_history=1,
                           action_history=1, goal_conditioned=True, num_diffusion_iters=1, local_map_size=20, max_batch=4096).eval()
kw = {"batch_size": int(os.environ.get("BATCH", "256")), "iteration_cap": int(os.environ.get("ITER_CAP", "4096"))}
sc.run_suite(sampler, total_runs=1, time_budget=1e9, planner_kwargs=kw)  # warm
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
table, secs = sc.run_suite(sampler, total_runs=2, time_budget=1e9, planner_kwargs=kw)
pr.disable()
torch.cuda.synchronize()
print(f"{len(table)} units in {time.time()-t0:.2f} s -> {len(table)/(time.time()-t0):.2f} units/s")
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
