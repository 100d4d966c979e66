"""MPPI baseline controller on the device kernels (config C4: K = 8192 rollouts).

PARITY UNPINNED.  The reference imports ``MPPI.mppi.MPPI`` from a module that is not in its
repository (run_scenarios_with_lidar_MPPI.py:10), so only the *interface* its driver uses is known:
``MPPI(maze_data, T, K, nx, nu)``, ``reset(start_state, goal_state)``, ``set_ref_path(path)``,
``is_done(state)``, ``step(state) -> (next_state, action, done | None)``, ``.env``,
``.reference_path`` (call sites :339-341, 392-402, 417-422, 442).  The arithmetic below is the
textbook MPPI update: K noisy control sequences are rolled out with the fused propagate+collide
kernel (the same dynamics and collision test the planner uses), a cost per rollout is formed, and
``dt_mppi_reduce`` computes w = softmax(-(c - min c)/lambda), u += sum_k w_k eps_k and the arg-min.
"""
from __future__ import annotations

import numpy as np
import torch

from .car_env import CarEnv
from .common.map_utils import _ctx_for


class MPPI:
    def __init__(self, maze_data, T=16, K=8192, nx=6, nu=2, lam=0.02, noise_sigma=(1.5, 0.8), lookahead=24,
                 collision_cost=1.0e4, effort_cost=1.0e-3):
        self.maze = np.asarray(maze_data, dtype=np.float32)
        self.T, self.K, self.nx, self.nu = int(T), int(K), int(nx), int(nu)
        self.lam = float(lam)
        self.lookahead = int(lookahead)
        self.collision_cost, self.effort_cost = float(collision_cost), float(effort_cost)
        self.env = CarEnv(maze_map=self.maze, collision_checking=False)
        self.ctx = _ctx_for(self.maze, 1.0)
        dev = self.ctx.device
        self.sigma = torch.tensor(noise_sigma, dtype=torch.float32, device=dev)
        self.u = torch.zeros((self.T, self.nu), dtype=torch.float32, device=dev)
        self.reference_path = None
        self._ref = None
        self.goal = None

    def reset(self, start_state=None, goal_state=None):
        self.u.zero_()
        if goal_state is not None:
            self.goal = np.asarray(goal_state, dtype=np.float64)
            self.env.goal = self.goal[:2].copy()
        if start_state is not None:
            self.env.set_state(np.asarray(start_state, dtype=np.float64))
        self.env.done = False
        self.env.terminated = False

    def set_ref_path(self, path):
        self.reference_path = np.asarray(path)
        self._ref = torch.as_tensor(self.reference_path[:, :2], dtype=torch.float32, device=self.ctx.device)

    def update_maze(self, new_maze):
        self.maze = np.asarray(new_maze, dtype=np.float32)
        self.env.maze_map = self.maze

    def is_done(self, state):
        return self.env.is_done(state)

    def rollout_costs(self, state, noise):
        """K rollouts of T steps from `state` with controls u + noise -> (cost (K,), target (2,)): ONE kernel
        (dt_mppi_rollout_cost: rollout, collision, look-ahead target, tracking / collision / effort cost)."""
        ctx = self.ctx = _ctx_for(self.maze, 1.0)
        s0 = torch.as_tensor(np.asarray(state, dtype=np.float32)).to(ctx.device, non_blocking=True)
        return ctx.mppi_rollout_cost(s0, self.u, noise, self._ref, self.lookahead, self.env.goal, self.collision_cost,
                                     self.effort_cost)

    def rollout_costs_reference(self, state, noise):
        """The same cost formed with the propagate kernel and elementwise torch ops (what rollout_costs fused): kept as
        the comparison for the kernel's test."""
        ctx = self.ctx = _ctx_for(self.maze, 1.0)
        s0 = torch.as_tensor(np.asarray(state, dtype=np.float32), device=ctx.device).expand(self.K, 6).contiguous()
        actions = (self.u[None] + noise).contiguous()
        res = ctx.propagate_collide(s0, actions, self.env.goal, want_traj=False, stop_on_collision=True)
        final = res["final"]
        cur = torch.as_tensor(np.asarray(state[:2], dtype=np.float32), device=ctx.device)
        near = int(torch.argmin(((self._ref - cur) ** 2).sum(1)))
        target = self._ref[min(near + self.lookahead, len(self._ref) - 1)]
        cost = ((final[:, :2] - target) ** 2).sum(1)
        cost = cost + self.collision_cost * (res["first_coll"] >= 0).float()
        cost = cost + self.effort_cost * (actions ** 2).sum((1, 2))
        return cost, target

    def step(self, curr_state):
        """-> (next_state, applied action, done) with done None on collision (driver convention :422-424).
        Device side of a tick: noise, rollout + cost, soft-min reduction, shift -- no host synchronisation until the
        applied action is read back."""
        if self._ref is None:
            raise ValueError("set_ref_path() must be called before step()")
        dev = self.ctx.device
        noise = torch.randn((self.K, self.T, self.nu), device=dev) * self.sigma
        cost, _ = self.rollout_costs(curr_state, noise)
        self.u, _, _ = self.ctx.mppi_reduce(cost, noise, self.lam, self.u)
        action = self.ctx.mppi_shift(self.u).cpu().numpy().astype(np.float64)
        self.env.set_state(np.asarray(curr_state, dtype=np.float64))
        self.env.collision_checking = True
        obs, _, terminated, _, info = self.env.step(action)
        self.env.collision_checking = False
        if terminated:
            self.env.terminated = False
            return obs, action, None
        return obs, action, bool(info["success"])
