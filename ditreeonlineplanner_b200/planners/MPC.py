"""Mirror of the reference's ``planners/MPC.py`` (``MPC_Planner``, :15-88): receding-horizon roll-out of the
sampler from the start state -- sample a chunk, propagate it, append a node, repeat; a collision or
``mpc_timeout`` chunks restart the chain from the start; the first chain that enters the goal disc wins.

The reference reads ``curr_node._state`` (:43), an attribute ``Node`` does not have, so its ``plan()`` raises
``AttributeError`` as shipped; this mirror implements the evident intent (``Node.state``).

``batch_size = B > 1`` runs B independent chains per device pass (local maps -> conditioning -> sampler ->
fused propagate + collide): a chain that collides or times out restarts from the start state in the next
pass, so all B slots always do useful work.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from ..common.map_utils import _ctx_for
from .base_planner import BasePlanner, Node


class MPC_Planner(BasePlanner):
    def __init__(self, start_state, goal_state, environment, sampler, **kwargs):
        super().__init__(start_state, goal_state, environment, sampler, **kwargs)
        self.mpc_timeout = 500          # restart from the start after this many chunks (MPC.py:18)
        self.batch_size = int(kwargs.get("batch_size", 1))
        self.iteration_cap = kwargs.get("iteration_cap", None)

    def reset(self):
        self.node_list.clear()
        self.node_list = [self.start_node]
        self.results = {"iterations": 0, "time": 0, "path": None, "actions": None, "number_of_nodes": 0}
        self.env.reset(options=self.options)

    def _local_map(self, state):
        n = int(self.local_map_size) if isinstance(self.local_map_size, (int, float)) else int(self.local_map_size[0])
        ctx = _ctx_for(self.maze, self.s_global)
        return ctx.local_map(torch.as_tensor(np.asarray(state[:3], dtype=np.float32)[None]), n, self.local_map_scale)

    def plan(self):
        if self.batch_size > 1:
            return self._plan_batched()
        start_time = time.time()
        goal = self.goal_state[:2]
        i = 0
        while time.time() - start_time <= self.time_budget:
            if self.iteration_cap is not None and i >= self.iteration_cap:
                break
            i += 1
            curr_node = self.start_node
            curr_state = curr_node.state
            states_sequence = curr_state[None, None, :] if curr_node.parent_states_seq is None else curr_node.parent_states_seq
            prev_actions = None
            j = 0
            while j < self.mpc_timeout and time.time() - start_time <= self.time_budget:
                j += 1
                sampled = self.sampler(states_sequence, prev_actions=prev_actions, goal=goal,
                                       local_map=self._local_map(curr_state))[0][:self.action_horizon]
                curr_state, done, actions_sequence, states_sequence = self.propagate_action_sequence_env(curr_state, sampled)
                prev_actions = sampled
                if curr_state is None or done is None:
                    break                                   # collision: start over
                node = Node(curr_state, actions_sequence, states_sequence, parent=curr_node)
                self.node_list.append(node)
                curr_node = node
                if done:
                    return self.handle_goal_reached(node, i, start_time)
        return self.handle_goal_not_reached(i, start_time)

    def _plan_batched(self):
        start_time = time.time()
        B = self.batch_size
        smp = self.sampler
        ctx = smp._context()
        _ctx_for(self.maze, self.s_global)
        n_map = int(self.local_map_size) if isinstance(self.local_map_size, (int, float)) else int(self.local_map_size[0])
        h, A, dev = self.action_horizon, smp.action_dim, ctx.device
        goal_xy = np.asarray(self.env.goal, dtype=np.float64)
        mean = torch.as_tensor(smp.metadata["Actions_mean"].astype(np.float32), device=dev)
        start = torch.as_tensor(np.asarray(self.start_node.state, dtype=np.float32), device=dev)
        goal_d = torch.as_tensor(np.asarray(self.goal_state[:2], dtype=np.float32), device=dev)
        states = start[None].repeat(B, 1)
        prev = mean[None].repeat(B, 1)
        age = torch.zeros(B, dtype=torch.int64, device=dev)
        chains = [[] for _ in range(B)]                     # per slot: list of (start state, actions, states) chunks
        passes = 0
        while time.time() - start_time <= self.time_budget:
            if self.iteration_cap is not None and passes * B >= self.iteration_cap:
                break
            passes += 1
            lm = ctx.local_map(states, n_map, self.local_map_scale, bf16_signed=True)
            cond = ctx.build_cond_car(states, prev, goal_d, smp.metadata, float(n_map))
            noise = torch.randn((B, smp.pred_horizon, A), device=dev)
            a = ctx.fm_sample(noise, cond, lm, smp.num_diffusion_iters, smp.metadata["Actions_mean"], smp.metadata["Actions_std"])
            res = ctx.propagate_collide(states, a, goal_xy, S=h, want_traj=True)
            coll = res["first_coll"] >= 0
            done = (res["done_step"] >= 0) & ~coll
            age = age + 1
            restart = coll | (age >= self.mpc_timeout)
            s0_h, a_h, t_h = states.cpu().numpy(), a[:, :h].cpu().numpy(), res["traj"].cpu().numpy()
            coll_h, done_h, restart_h = coll.cpu().numpy(), done.cpu().numpy(), restart.cpu().numpy()
            steps_h = np.where(done_h, res["done_step"].cpu().numpy() + 1, h)
            for b in range(B):
                if coll_h[b]:
                    chains[b] = []
                    continue
                chains[b].append((s0_h[b], a_h[b, :steps_h[b]], t_h[b, :steps_h[b]]))
                if done_h[b]:
                    node = self.start_node
                    for s0, acts, st in chains[b]:          # materialise the winning chain as nodes
                        node = Node(st[-1].astype(np.float64), acts.astype(np.float64),
                                    np.concatenate([s0[None], st]).astype(np.float64)[None], parent=node)
                        self.node_list.append(node)
                    return self.handle_goal_reached(node, passes * B, start_time)
                if restart_h[b]:
                    chains[b] = []
            states = torch.where(restart[:, None], start[None], res["final"])
            prev = torch.where(restart[:, None], mean[None], a[:, h - 1])
            age = torch.where(restart, torch.zeros_like(age), age)
        return self.handle_goal_not_reached(passes * B, start_time)
