"""Drop-in mirror of the reference's ``car_env.py::CarEnv`` for the planner-facing surface
(constructor :26, reset :206, step :240-282, set_state :306, state :169, is_done :176,
cell_rowcol_to_xy / cell_xy_to_rowcol :189-201, maze_map property :112-128), with the dynamics
(_update_state :356-396), the goal test (_check_done :341-354) and the optional in-env collision
test executed by the fused propagate+collide kernel.

The scalar ``step`` costs one kernel launch; batched work goes through
``ditreeonlineplanner_b200.expansion.TreeExpander`` / ``BatchedCarEnv.step``.
run_type >= 2 (probability-map state sampling): ``prior`` is the normalised Euclidean distance transform of
the free cells (``dt_edt_prior``), ``prob_map`` the prior (run_type 2) or its log-blend with a Gaussian along
the robot -> goal line (run_type >= 3, ``dt_prob_map``), recomputed where the reference recomputes them
(constructor, ``maze_map`` setter, ``update_prob_map_by_loc``; car_env.py:98-137).
"""
from __future__ import annotations

import types
from typing import Dict, Optional

import numpy as np
import torch

from .common.map_utils import _ctx_for, invalidate_staged_map


def bicycle_model():
    """Constants of the reference's bicycle model (car_env.py:493-625): state/control sizes and
    the input bounds the env clips actions to."""
    model = types.SimpleNamespace(
        name="CartesianBicycleModel", nx=6, nu=2, x0=np.zeros(6),
        throttle_min=-1.0, throttle_max=1.0, delta_min=-0.40, delta_max=0.40,
        ddelta_min=-2.0, ddelta_max=2.0, dthrottle_min=-10, dthrottle_max=10,
        params=types.SimpleNamespace(m=0.043, C1=0.5, C2=15.5, Cm1=0.28, Cm2=0.05, Cr0=0.011, Cr2=0.006))
    constraint = types.SimpleNamespace(alat_min=-4, alat_max=4, along_min=-4, along_max=4)
    return model, constraint


class _Box:
    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        self.shape = self.low.shape
        self.dtype = np.float32


class CarEnv:
    metadata: dict = {}

    def __init__(self, lidar2dsim=None, dt=0.02, drone_radius=0.1, maze_map=None, collision_checking=True, run_type=0):
        if maze_map is None:
            raise ValueError("maze_map is required")
        if lidar2dsim is None:
            from .lidar_sim.lidar_2d_sim import Lidar2DSim
            lidar2dsim = Lidar2DSim()
        self.lidar2dsim = lidar2dsim
        self.dt = 1.0 / 50.0
        self.current_step = 0
        self.collision_checking = collision_checking
        self.ball_radius = drone_radius
        self.model, self.constraints = bicycle_model()
        self.state_dim, self.action_dim = 6, 2
        self.car_length, self.car_width = 0.35, 0.2
        self.m, self.C1, self.C2, self.Cm1, self.Cm2, self.Cr0, self.Cr2 = 0.043, 0.5, 15.5, 0.28, 0.05, 0.011, 0.006
        self.action_space = _Box([self.model.dthrottle_min, self.model.ddelta_min],
                                 [self.model.dthrottle_max, self.model.ddelta_max])
        self.observation_space = _Box(np.full(6, -np.inf), np.full(6, np.inf))
        self._state = self.model.x0.astype(np.float64)
        self._maze_map = np.asarray(maze_map)
        self._maze_height = 1
        self._maze_size_scaling = 1
        self._map_length = len(maze_map)
        self._map_width = len(maze_map[0])
        self._x_map_center = self._map_width / 2 * self._maze_size_scaling
        self._y_map_center = self._map_length / 2 * self._maze_size_scaling
        self.goal = np.array([0, 0])
        self.done = False
        self.terminated = False
        self.run_type = run_type
        self.gaussian_pdf = None
        self._refresh_prior()
        if self.run_type < 2:
            self.prob_map = np.zeros_like(self._maze_map.copy())   # unused by the original / +reference runs
        elif self.run_type == 2:
            self.prob_map = self.prior
        else:  # the constructor passes (row, col) un-swapped (car_env.py:107-109); the other call sites swap
            self._blend(self.cell_xy_to_rowcol(self.state[:2]), self.cell_xy_to_rowcol(self.goal[:2]))

    # ---- map ------------------------------------------------------------------------------
    @property
    def maze_map(self):
        return self._maze_map

    @maze_map.setter
    def maze_map(self, new_maze_map):
        self._maze_map = np.asarray(new_maze_map)
        invalidate_staged_map()
        self._refresh_prior()
        if self.run_type == 2:
            self.prob_map = self.prior
        elif self.run_type >= 3:
            self.update_prob_map_by_loc()

    # ---- probability-map sampler (car_env.py:98-137) ------------------------------------------
    def _refresh_prior(self):
        """prior = distance_transform_edt(1 - maze) / sum, on the device (skipped while nothing samples it)."""
        if self.run_type < 2:
            self.prior = None
            return
        self.prior = _ctx_for(self._maze_map, 1.0).edt_prior().cpu().numpy()

    def _blend(self, robot, goal):
        from .prob_sampling_utils import blended_prob_map
        if tuple(self.prior.shape) != (20, 20):   # gaussian_map's default size (prob_sampling_utils.py:48)
            raise ValueError(f"operands could not be broadcast together with shapes {self.prior.shape} (20,20)")
        self.prob_map, self.gaussian_pdf = blended_prob_map(self.prior, robot, goal)

    def update_prob_map_by_loc(self):
        self._blend(self.cell_xy_to_rowcol(self.state[:2])[::-1], self.cell_xy_to_rowcol(self.goal[:2])[::-1])

    @property
    def maze_size_scaling(self):
        return self._maze_size_scaling

    @property
    def maze_height(self):
        return self._maze_height

    @property
    def x_map_center(self):
        return self._x_map_center

    @property
    def y_map_center(self):
        return self._y_map_center

    # ---- state ----------------------------------------------------------------------------
    @property
    def state(self):
        return self._get_obs()

    def _get_obs(self):
        return np.array(self._state, copy=True)

    def set_state(self, state):
        self._state = state

    def reset_done(self):
        self.done = False

    def _goal_reached(self, xy):
        return bool(np.linalg.norm(np.asarray(xy[:2], dtype=np.float64) - self.goal) < 0.5)

    def is_done(self, curr_state):
        return self._goal_reached(curr_state)

    def cell_rowcol_to_xy(self, rowcol_pos):
        x = (rowcol_pos[1] + 0.5) * self.maze_size_scaling - self.x_map_center
        y = self.y_map_center - (rowcol_pos[0] + 0.5) * self.maze_size_scaling
        return np.array([x, y])

    def cell_xy_to_rowcol(self, xy_pos, floor_enable=True):
        i = (self.y_map_center - xy_pos[1]) / self.maze_size_scaling
        j = (xy_pos[0] + self.x_map_center) / self.maze_size_scaling
        ret = np.array([i, j])
        return np.floor(ret) if floor_enable else ret

    def reset(self, *, seed: Optional[int] = None, options: Optional[Dict[str, Optional[np.ndarray]]] = None, **kwargs):
        self._state = np.zeros(6, dtype=np.float32)
        if options is not None:
            if options.get("goal_cell") is not None:
                self.goal = self.cell_rowcol_to_xy(options["goal_cell"])
            if options.get("reset_cell") is not None:
                self._state[0:2] = self.cell_rowcol_to_xy(options["reset_cell"])
            if options.get("reset_deg") is not None:
                self._state[2] = np.deg2rad(options["reset_deg"])
        self.current_step = 0
        self.done = False
        self.terminated = False
        return self._get_obs(), None

    # ---- dynamics -------------------------------------------------------------------------
    def step(self, action):
        """-> (obs, reward, terminated, False, info{collision, goal, success}) (car_env.py:240-282)."""
        collision = False
        if not self.done and not self.terminated:
            ctx = _ctx_for(self._maze_map, 1.0)
            s0 = torch.as_tensor(np.asarray(self._state, dtype=np.float32)[None])
            act = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, 1, 2))
            res = ctx.propagate_collide(s0, act, self.goal, want_traj=False, stop_on_collision=False)
            self._state = res["final"][0].cpu().numpy().astype(np.float64)
            reward = 0
            self.current_step += 1
            self.done = int(res["done_step"][0]) >= 0
            if self.collision_checking:
                collision = int(res["first_coll"][0]) >= 0
            if collision:
                reward = -1.0
                self.terminated = True
        else:
            reward = 0.0
        info = {"collision": collision, "goal": self.goal, "success": self.done}
        return self._get_obs(), reward, self.terminated, False, info

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the hot path")


class BatchedCarEnv:
    """B independent car environments on the device (the vectorised env the reference builds with
    SB3's DummyVecEnv in rollout_manager.py:612-664, without the per-env Python loop)."""

    def __init__(self, maze_map, goals_xy, device=None):
        self.maze = np.asarray(maze_map, dtype=np.float32)
        self.ctx = _ctx_for(self.maze, 1.0, device)
        self.goals = np.asarray(goals_xy, dtype=np.float32)

    def rollout(self, states, actions, steps=None, want_traj=True, stop_on_collision=True):
        """All candidates share goals[0] when a single goal is given.  states (B,6), actions (B,T,2)."""
        self.ctx = _ctx_for(self.maze, 1.0)
        g = self.goals if self.goals.ndim == 1 else self.goals[0]
        return self.ctx.propagate_collide(states, actions, g, S=steps, want_traj=want_traj,
                                          stop_on_collision=stop_on_collision)
