#!/usr/bin/env python
"""cProfile of the reference-order B = 1 planning loop (bench.py's C1 leg): where the host time per iteration goes."""
import cProfile, os, pstats, random, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ditreeonlineplanner_b200 import load_scenarios
from ditreeonlineplanner_b200 import scenarios as sc
from ditreeonlineplanner_b200.car_env import CarEnv
from ditreeonlineplanner_b200.data import load_maze
from ditreeonlineplanner_b200.planners.RRT import RRT_Planner
from ditreeonlineplanner_b200.policies.fm_policy import DiffusionSampler
from ditreeonlineplanner_b200.weights import UNET_DIMS, random_init

dims = UNET_DIMS["large"]
sd = random_init(seed=0, input_dim=2, cond_dim=7, emb_dim=400, down_dims=dims)
sampler = DiffusionSampler(sd, None, "carmaze", policy="flow_matching", pred_horizon=64, action_dim=2, obs_history=1,
                           action_history=1, goal_conditioned=True, num_diffusion_iters=1, local_map_size=20,
                           max_batch=64).eval()
row = load_scenarios("test_scenarios_car")[0]
maze = load_maze(row["maze_name"])
env = CarEnv(maze_map=maze, collision_checking=False)
start, goal_s = sc.scenario_states(row, env)
pl = RRT_Planner(start, goal_s, env_id="carmaze", environment=env, sampler=sampler, prediction_type="actions",
                 action_horizon=8, local_map_size=20, local_map_scale=0.2, global_map_scale=1.0,
                 goal_conditioning_bias=0.85, prop_duration=[64], time_budget=1e9, max_iter=300, iteration_cap=100)
prof = cProfile.Profile()
for mode in ("warm", "plain", "profiled"):
    pl.iteration_cap = 40 if mode == "warm" else 1000
    torch.manual_seed(42); np.random.seed(42); random.seed(42)
    pl.reset()
    t0 = time.perf_counter()
    if mode == "profiled":
        prof.enable(); pl.plan(); prof.disable()
    else:
        pl.plan()
    torch.cuda.synchronize()
    print("%s: %.1f iterations/s" % (mode, pl.results["iterations"] / (time.perf_counter() - t0)), flush=True)
pstats.Stats(prof).sort_stats("cumulative").print_stats(45)
