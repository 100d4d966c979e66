"""SB3-free mirror of the reference's batched validation rollout
(``rollout_manager.py::rollout`` :545-838) for the car environment: B = scenarios x
``envs_per_scenario`` environments advanced together -- one batched sampler call per chunk
(local maps + conditioning + denoiser on the device) and one fused propagate+collide launch per
scenario group instead of SB3's ``DummyVecEnv`` stepping B Python envs one by one.

Same signature and the same per-scenario result dicts (keys of rollout_manager.py:825-836).
Semantics kept: the previous ``action_history`` actions condition the next call, a collision or
reaching the goal ends an environment (its trajectory row is then held constant, :765), and
``step_to_completion`` is the first step at which the goal was reached (-1 if never).
Documented deviation: SB3's vec-env silently *resets* an environment that terminates; here a
terminated environment simply stays frozen, which is what the reference's own bookkeeping assumes.
"""
from __future__ import annotations

import numpy as np
import torch

from .common.map_utils import _ctx_for
from .data import load_maze, load_scenarios
from .policies.fm_policy import DiffusionSampler


def rollout(env_id, policy, ema_noise_pred_net, noise_scheduler, max_episode_steps=250, render_mode="rgb_array",
            num_diffusion_iters=100, prediction_type="actions", obs_history=1, action_history=1,
            position_conditioned=False, goal_conditioned=True, local_map_conditioned=True, local_map_size=10, scale=0.2,
            pred_horizon=16, action_horizon=8, envs_per_scenario=32, render=False, scenarios=None):
    if "car" not in env_id.lower():
        raise NotImplementedError("only the car environment has device dynamics (ant / point need MuJoCo)")
    if prediction_type != "actions":
        raise NotImplementedError("prediction_type='observations' (PD tracking) is outside the hot path")
    sampler = DiffusionSampler(ema_noise_pred_net, noise_scheduler, env_id, policy, pred_horizon, 2, prediction_type,
                               obs_history, action_history, num_diffusion_iters, local_map_size=local_map_size,
                               max_batch=4096).eval()
    rows = scenarios if scenarios is not None else load_scenarios("validation_scenarios_car")
    E = envs_per_scenario
    mazes = [load_maze(r["maze_name"]) for r in rows]
    n_sc = len(rows)
    B = n_sc * E
    start_rc = [(int(r["start_row"]), int(r["start_col"])) for r in rows]
    goal_rc = [(int(r["goal_row"]), int(r["goal_col"])) for r in rows]

    def rc_to_xy(rc, maze):
        R, C = maze.shape
        return np.array([(rc[1] + 0.5) - C / 2, R / 2 - (rc[0] + 0.5)])
    start_xy = [rc_to_xy(start_rc[i], mazes[i]) for i in range(n_sc)]
    goal_xy = [rc_to_xy(goal_rc[i], mazes[i]) for i in range(n_sc)]
    ctx = sampler._context()
    dev = ctx.device
    obs = np.zeros((B, 6), dtype=np.float32)
    goal = np.zeros((B, 2), dtype=np.float32)
    for i, r in enumerate(rows):
        obs[i * E:(i + 1) * E, :2] = start_xy[i]
        obs[i * E:(i + 1) * E, 2] = np.deg2rad(float(r["start_deg"]))
        goal[i * E:(i + 1) * E] = goal_xy[i]
    state = torch.as_tensor(obs, device=dev)
    goal_d = torch.as_tensor(goal, device=dev)
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    collision_count = torch.zeros(B, device=dev)
    step_to_completion = torch.full((B,), float("inf"), device=dev)
    best_dist = torch.full((B,), float("inf"), device=dev)
    traj = torch.zeros((B, max_episode_steps + 1, 8), device=dev)
    prev_action = None
    h = action_horizon
    curr_step = 0
    while curr_step < max_episode_steps and not bool(done.all()):
        n = min(h, max_episode_steps - curr_step)
        # batched sampler call: per-maze local maps, per-environment goals
        lms = []
        for i in range(n_sc):
            _ctx_for(mazes[i], 1.0)
            lms.append(ctx.local_map(state[i * E:(i + 1) * E], int(local_map_size), scale, bf16_signed=True))
        lm = torch.cat(lms)
        cond = ctx.build_cond_car(state, prev_action, goal_d, sampler.metadata, float(local_map_size))
        noise = torch.randn((B, pred_horizon, 2), device=dev)
        actions = ctx.fm_sample(noise, cond, lm, num_diffusion_iters, sampler.metadata["Actions_mean"],
                                sampler.metadata["Actions_std"])
        new_state = state.clone()
        for i in range(n_sc):
            sl = slice(i * E, (i + 1) * E)
            _ctx_for(mazes[i], 1.0)
            res = ctx.propagate_collide(state[sl], actions[sl], goal_xy[i], S=n, want_traj=True, stop_on_collision=True)
            tr, first, dn = res["traj"], res["first_coll"], res["done_step"]
            live = ~done[sl]
            steps_run = torch.where(first >= 0, first + 1, torch.where(dn >= 0, dn + 1, torch.full_like(first, n)))
            t_idx = torch.arange(n, device=dev)[None, :]
            ran = (t_idx < steps_run[:, None]) & live[:, None]  # steps the environment actually took
            # trajectory rows: (state before the step, action of the step); finished envs hold their last row
            before = torch.cat([state[sl][:, None, :], tr[:, :-1, :]], dim=1) if n > 1 else state[sl][:, None, :]
            rows_new = torch.cat([before, actions[sl][:, :n, :]], dim=-1)
            seg = traj[sl, curr_step:curr_step + n]
            seg[ran] = rows_new[ran]
            d = torch.linalg.norm(tr[..., :2] - goal_d[sl][:, None, :], dim=-1)
            d = torch.where(ran, d, torch.full_like(d, float("inf")))
            best_dist[sl] = torch.minimum(best_dist[sl], d.min(dim=1).values)
            succ = live & (dn >= 0)
            step_to_completion[sl] = torch.where(succ, torch.minimum(step_to_completion[sl], (curr_step + dn).float()),
                                                 step_to_completion[sl])
            collision_count[sl] += (live & (first >= 0)).float()
            new_state[sl] = torch.where(live[:, None], res["final"], state[sl])
            done[sl] = done[sl] | (first >= 0) | (dn >= 0)
        # hold the last written row for steps an environment did not take
        for t in range(curr_step, curr_step + n):
            if t > 0:
                empty = (traj[:, t].abs().sum(-1) == 0)
                traj[empty, t] = traj[empty, t - 1]
        state = new_state
        prev_action = actions[:, n - 1, :].contiguous()
        curr_step += n
    traj[:, curr_step] = torch.cat([state, torch.zeros((B, 2), device=dev)], dim=-1)
    if curr_step > 0:
        traj[done, curr_step] = traj[done, curr_step - 1]
    stc = torch.where(torch.isinf(step_to_completion), torch.full_like(step_to_completion, -1.0), step_to_completion)
    best_dist, stc, traj, collision_count = (t.cpu().numpy() for t in (best_dist, stc, traj, collision_count))
    results = []
    for i, r in enumerate(rows):
        sl = slice(i * E, (i + 1) * E)
        results.append({"scenario_name": r["scenario_name"], "maze": mazes[i], "start_rowcol": start_rc[i],
                        "goal_rowcol": goal_rc[i], "start_position": start_xy[i], "goal_position": goal_xy[i],
                        "best_dist": best_dist[sl], "step_to_completion": stc[sl], "trajectory": traj[sl],
                        "collision_count": collision_count[sl]})
    return results, []
