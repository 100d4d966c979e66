"""Drop-in mirror of the reference's ``planners/base_planner.py`` for the car / ant hot path:
``Node`` (:24-34), ``BasePlanner`` constructor (:38-135), ``check_collision`` (:145-155),
``random_node_sample`` (:162-207), ``propagate_action_sequence_env`` (:257-320),
``generate_final_path_env`` (:342-363), goal bookkeeping (:231-255).  Collision checks and
propagation run on the device through the C ABI."""
from __future__ import annotations

import abc
import random
import time
from collections import deque

import numpy as np
import torch

from ..common.map_utils import _ctx_for, is_colliding_ant, is_colliding_car


class Node:
    __slots__ = ("state", "parent_action_seq", "parent_states_seq", "parent", "cached_actions", "num_visit", "index")

    def __init__(self, state, parent_action_seq=None, parent_states_seq=None, parent=None):
        self.state = state
        self.parent_action_seq = parent_action_seq  # (n, act_dim) actions of the edge into this node
        self.parent_states_seq = parent_states_seq  # (1, n, obs_dim) states along that edge
        self.parent = parent
        self.cached_actions = deque([])
        self.num_visit = 0
        self.index = -1

    def __repr__(self):
        return f"Node(State={self.state}, nVisits={self.num_visit})"


class BasePlanner(abc.ABC):
    def __init__(self, start_state, goal_state, environment, sampler, action_horizon=8, local_map_size=(10, 10),
                 local_map_scale=0.2, global_map_scale=1.0, env_id="pushT", time_budget=10, **kwargs):
        if environment is None:
            raise ValueError("Environment is not defined.")
        self.env = environment
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        if isinstance(sampler, torch.nn.Module):
            self.sampler = sampler.to(self.device)
        elif sampler is not None:
            self.sampler = sampler
        self.action_horizon = action_horizon
        self.local_map_size = local_map_size
        self.local_map_scale = local_map_scale
        self.s_global = global_map_scale
        self.env_id = env_id
        self.time_budget = time_budget
        self.start_node = Node(start_state)
        self.goal_state = goal_state
        self.node_list = [self.start_node]
        self.results = {"iterations": 0, "time": 0, "path": None, "actions": None, "number_of_nodes": 0}
        self.render = kwargs.get("render", False)
        self.verbose = kwargs.get("verbose", False)
        self.env_dt = self.env.dt if hasattr(self.env, "dt") else 0.1
        if "car" in env_id.lower():
            start = self.env.cell_xy_to_rowcol(start_state[:2])
            goal = self.env.cell_xy_to_rowcol(goal_state[:2])
            self.options = {"reset_cell": start, "reset_deg": np.rad2deg(start_state[2]), "goal_cell": goal}
            self.env.reset(options=self.options)
            self.max_v = 5
            self.x_center = self.env.x_map_center
            self.y_center = self.env.y_map_center
            self.map_width = len(self.env.maze_map[0])
            self.map_length = len(self.env.maze_map)
            self.maze = np.float32(self.env.maze_map)
        else:
            raise NotImplementedError(f"env_id {env_id!r}: only the car environment has a device dynamics model "
                                      "(the ant's MuJoCo dynamics is outside the hot path)")
        self.save_bad_edges = False
        self.failed_node_list = []
        self._debug = kwargs.get("debug", False)
        self._scenario_num = str(kwargs.get("scenario_num", "999"))
        self._scenario_name = str(kwargs.get("scenario_name", "test"))
        self.scenario_iter_num = str(kwargs.get("iter_num", "0"))
        self.save_path = str(kwargs.get("root_folder", "benchmark_results"))

    @property
    def scenario_iter_folder_name(self):
        return f"Iter_{self.scenario_iter_num}"

    @abc.abstractmethod
    def plan(self):
        pass

    @abc.abstractmethod
    def reset(self):
        pass

    def check_collision(self, state=None):
        if "car" in self.env_id.lower():
            return is_colliding_car(state, self.maze)
        if "ant" in self.env_id.lower():
            return is_colliding_ant(state, self.maze, 1.2, self.s_global)
        raise NotImplementedError(self.env_id)

    def sample_row_col_from_probability_map(self):
        """np.random.choice(size, size=1, p=prob_map.ravel()) (base_planner.py:157-160): the one uniform
        variate RandomState.choice consumes is drawn here, the inverse-CDF search runs on the device."""
        pm = np.asarray(self.env.prob_map, dtype=np.float64)
        s = pm.sum()
        if not np.all(pm >= 0) or abs(s - 1.0) > 1.5e-8:   # NumPy's own checks (sqrt(eps) tolerance)
            raise ValueError("probabilities do not sum to 1" if np.all(pm >= 0) else "probabilities are not non-negative")
        u = np.random.random_sample(1)
        flat = _ctx_for(self.maze, self.s_global).sample_cells(pm, u).cpu().numpy()
        row, col = np.unravel_index(flat, pm.shape)
        return row[np.newaxis], col[np.newaxis]

    def random_node_sample(self, batch_size=1):
        """Same RNG consumption as the reference (python `random`, then the cell draw or two uniform draws,
        then four np.random.uniform draws)."""
        if random.random() > self.goal_sample_rate:
            if getattr(self, "run_type", 0) >= 2:
                rows, cols = self.sample_row_col_from_probability_map()
                x, y = self.env.cell_rowcol_to_xy(np.array([rows[0], cols[0]]))
                x, y = x[np.newaxis], y[np.newaxis]
            else:
                x = np.random.uniform(-self.map_width / 2, self.map_width / 2, size=(batch_size, 1))
                y = np.random.uniform(-self.map_length / 2, self.map_length / 2, size=(batch_size, 1))
            theta = np.random.uniform(-np.pi, np.pi, size=(batch_size, 1))
            v = np.random.uniform(-self.max_v, self.max_v, size=(batch_size, 1))
            throttle = np.random.uniform(-1, 1, size=(batch_size, 1))
            steer = np.random.uniform(-0.40, 0.40, size=(batch_size, 1))
            return np.concatenate((x, y, theta, v, throttle, steer), axis=1)
        sample = np.zeros((batch_size, self.start_node.state.shape[0]))
        sample[:] = self.goal_state
        return sample

    def dist_to_goal(self, state):
        return np.linalg.norm(state[:2] - self.goal_state[:2])

    def handle_goal_reached(self, node, iterations, start_time):
        self.results["time"] = time.time() - start_time
        path, actions = self.generate_final_path_env(node)
        self.results["iterations"] = iterations
        self.results["path"] = path
        self.results["path_time"] = len(path) * self.env_dt
        self.results["actions"] = actions
        self.results["number_of_nodes"] = len(self.node_list)
        if self.verbose:
            print(f" Goal reached in {iterations} iterations.")
        return path, actions

    def handle_goal_not_reached(self, iterations, start_time):
        self.results["time"] = time.time() - start_time
        self.results["iterations"] = iterations
        self.results["number_of_nodes"] = len(self.node_list)
        if self.verbose:
            print(f" Goal not reached in {iterations} iterations.")
        return None, None

    # ---- propagation ------------------------------------------------------------------------
    def propagate_action_sequence_env(self, state, action_sequence):
        """-> (obs, done in {True, False, None = collision}, actions (<= h, A), states (1, <= h + 1, D)).
        One fused kernel launch (B = 1) instead of h Python env steps + h collision checks."""
        if action_sequence is None:
            raise ValueError("Action sequence is None.")
        h = self.action_horizon
        state = np.asarray(state, dtype=np.float64)
        self.env.set_state(state)
        n = len(action_sequence[:h])
        states_sequence = np.zeros((h + 1, state.shape[0]))
        states_sequence[0] = state
        if self.env.done or self.env.terminated or n == 0:
            return self._propagate_latched(state, action_sequence, states_sequence, n)
        ctx = _ctx_for(self.maze, 1.0)
        # one host->device copy (state | actions) and one device->host copy (trajectory | final state | flags): this runs
        # once per iteration of the reference's B = 1 loop, where every extra copy or synchronisation is ~10 us
        na = 2 * n + (-2 * n) % 4                    # actions first: both parts stay 16-byte aligned on the device
        flat = np.zeros(na + 6, dtype=np.float32)
        flat[:2 * n] = np.asarray(action_sequence[:n], dtype=np.float32).reshape(-1)
        flat[na:] = state
        dev = torch.from_numpy(flat).to(ctx.device)
        res = ctx.propagate_collide(dev[na:].view(1, 6), dev[:2 * n].view(1, n, 2), self.env.goal, want_traj=True,
                                    stop_on_collision=True, packed_out=True)
        host = res["blob"].cpu().numpy()
        pitch = res["pitch"]
        first, done_step = int(host[pitch + 6: pitch + 7].view(np.int32)[0]), int(host[pitch + 7: pitch + 8].view(np.int32)[0])
        traj = host[: n * 6].reshape(n, 6).astype(np.float64)
        if self.maze.shape[0] > self.maze.shape[1]:
            ctx.sync_status()   # IndexError like the reference's clipped diagonal lookup: only tall maps can raise it
        obs = host[pitch: pitch + 6].astype(np.float64)
        self.env.set_state(obs.copy())
        self.env.current_step += (first + 1) if first >= 0 else ((done_step + 1) if done_step >= 0 else n)
        if first >= 0:
            states_sequence[1:first + 2] = traj[:first + 1]
            # the reference's env.step ran its goal test (and, with collision_checking, its own collision test)
            # on the colliding state BEFORE the planner saw the collision (car_env.py:264-272): both latches
            # freeze every later propagation of this env
            if self.env._goal_reached(obs[:2]):
                self.env.done = True
            if self.env.collision_checking:
                self.env.terminated = True
            return obs, None, action_sequence[:first], states_sequence[:first][None, :]
        done = False
        if done_step >= 0:
            states_sequence[1:done_step + 2] = traj[:done_step + 1]
            action_sequence[done_step + 1:] = 0
            self.env.done = True
            done = True
        else:
            states_sequence[1:n + 1] = traj[:n]
        action_sequence = action_sequence[:h]
        states_sequence = states_sequence[:len(action_sequence) + 1]
        return obs, done, action_sequence[:h], states_sequence[None, :]

    def _propagate_latched(self, state, action_sequence, states_sequence, n):
        """The env froze after reaching the goal / colliding (car_env.py:254,274-275): every step
        returns the same state, so only the flags matter."""
        h = self.action_horizon
        obs = state
        done = False
        for i in range(n):
            done = bool(self.env.done)
            states_sequence[i + 1] = obs
            if self.check_collision(obs):
                return obs, None, action_sequence[:i], states_sequence[:i][None, :]
            if done:
                action_sequence[i + 1:] = 0
                break
        action_sequence = action_sequence[:h]
        states_sequence = states_sequence[:len(action_sequence) + 1]
        return obs, done, action_sequence[:h], states_sequence[None, :]

    def generate_final_path_env(self, final_node):
        """Back-trace: states of every edge followed by its end node, actions of every edge."""
        chain = []
        node = final_node
        while node is not None:
            chain.append(node)
            node = node.parent
        path, acts = [], []
        for nd in reversed(chain):
            if nd.parent_states_seq is not None:
                path.extend(list(nd.parent_states_seq[0]))
            path.append(nd.state)
            if nd.parent_action_seq is not None:
                acts.extend(list(nd.parent_action_seq))
        return (np.array(path, dtype=np.float32) if path else None), (np.array(acts, dtype=np.float32) if acts else None)
