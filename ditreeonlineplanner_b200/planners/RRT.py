"""Drop-in mirror of the reference's ``planners/RRT.py::RRT_Planner`` (constructor :19-31, reset
:33-47, nearest_node[_batch] :49-55, update_maze :57-59, check_obstacle_ahead :61-81, plan :113-257)
whose expansion runs on the B200:

* ``batch_size = 1`` (default) follows the reference loop statement by statement -- same RNG
  consumption order, one node inserted per iteration -- with the nearest-neighbour
  query, local map, sampler, propagation and collision checks executed by the device kernels.
* ``batch_size = B > 1`` expands B sampled states per round in one batched device pass per chunk
  (``TreeExpander``): every candidate edge is computed exactly as in the B = 1 loop, but the B
  samples of a round see the tree as it was at the start of the round (documented deviation).

The wall-clock budget is kept (``time.time()``); ``iteration_cap`` (kwarg, optional) bounds the
number of sampler calls for deterministic runs.
"""
from __future__ import annotations

import random
import time

import numpy as np
import torch

from ..common.map_utils import _ctx_for
from .base_planner import BasePlanner, Node


# host seconds spent waiting for device passes / passes booked by the continuous planner in this process
# (bench.py reports them per rank next to the suite throughput: device-bound vs host-bound at a glance)
PASS_STATS = {"gpu_wait_s": 0.0, "passes": 0}

class _DeviceTree:
    """SoA mirror of the node positions in HBM (x[], y[]) for the nearest-neighbour kernel; replaces
    the KD-tree the reference rebuilds from scratch after every insertion (RRT.py:207)."""

    def __init__(self, device, capacity=1024):
        self.device = device
        self.cap = capacity
        self.x = torch.empty(capacity, dtype=torch.float32, device=device)
        self.y = torch.empty(capacity, dtype=torch.float32, device=device)
        self.n = 0

    def clear(self):
        self.n = 0

    def append(self, xy_rows):
        xy = torch.as_tensor(np.asarray(xy_rows, dtype=np.float32).reshape(-1, 2), device=self.device)
        k = xy.shape[0]
        if self.n + k > self.cap:
            while self.n + k > self.cap:
                self.cap *= 2
            nx = torch.empty(self.cap, dtype=torch.float32, device=self.device)
            ny = torch.empty(self.cap, dtype=torch.float32, device=self.device)
            nx[:self.n], ny[:self.n] = self.x[:self.n], self.y[:self.n]
            self.x, self.y = nx, ny
        self.x[self.n:self.n + k] = xy[:, 0]
        self.y[self.n:self.n + k] = xy[:, 1]
        self.n += k


class RRT_Planner(BasePlanner):
    def __init__(self, start_state, goal_state, environment, sampler, **kwargs):
        super().__init__(start_state, goal_state, environment, sampler, **kwargs)
        self.kd_tree_dim = 2
        self.goal_sample_rate = 0.15
        self.goal_conditioning_bias = kwargs.get("goal_conditioning_bias", 0.85)
        self.prop_duration_schedule = kwargs.get("prop_duration", [64])
        self.offline_time_budget = kwargs.get("offline_time_budget", 60)
        self.plan_count = 0
        self.init_main_path = None
        self.run_type = kwargs.get("run_type", 0)
        self.env.run_type = self.run_type
        self.batch_size = int(kwargs.get("batch_size", 1))
        self.iteration_cap = kwargs.get("iteration_cap", None)
        # batched expansion flavour: "continuous" refills a slot as soon as its edge ends (every slot does
        # useful work in every device pass); "rounds" expands one batch of edges to completion at a time
        self.batch_mode = kwargs.get("batch_mode", "continuous")
        if self.batch_mode not in ("continuous", "rounds", "device"):
            raise ValueError("batch_mode must be 'continuous', 'rounds' or 'device'")
        self._ctx = _ctx_for(self.maze, 1.0)
        self._tree = _DeviceTree(self._ctx.device)
        self._tree.append(np.asarray(start_state[:2]))
        self.start_node.index = 0

    # ---- bookkeeping ------------------------------------------------------------------------
    def reset(self, start_state: np.ndarray = None, goal_state: np.ndarray = None, reset_main_path: bool = False):
        if reset_main_path:
            self.init_main_path = None
        if start_state is not None:
            self.start_node = Node(start_state)
            self.goal_state = goal_state
            self.options["reset_cell"] = self.env.cell_xy_to_rowcol(start_state[:2])
            self.options["reset_deg"] = np.rad2deg(start_state[2])
            self.options["goal_cell"] = self.env.cell_xy_to_rowcol(goal_state[:2])
        self.node_list = [self.start_node]
        self.start_node.index = 0
        self.failed_node_list = []
        self._tree.clear()
        self._tree.append(np.asarray(self.start_node.state[:2]))
        self.results = {"iterations": 0, "time": 0, "path": None, "actions": None, "number_of_nodes": 0}
        self.env.reset(options=self.options)

    def _insert(self, node):
        node.index = len(self.node_list)
        self.node_list.append(node)
        self._tree.append(np.asarray(node.state[:2]))

    def _insert_many(self, nodes):
        """One device append for a whole round of new nodes (the batched planner)."""
        if not nodes:
            return
        for node in nodes:
            node.index = len(self.node_list)
            self.node_list.append(node)
        self._tree.append(np.stack([np.asarray(n.state[:2]) for n in nodes]))

    def nearest_node(self, sample):
        idx = self._ctx.nearest(self._tree.x[:self._tree.n], self._tree.y[:self._tree.n],
                                torch.as_tensor(np.asarray(sample, dtype=np.float32)[:, :2]))
        return self.node_list[int(idx[0])]

    def nearest_node_batch(self, samples):
        idx = self._ctx.nearest(self._tree.x[:self._tree.n], self._tree.y[:self._tree.n],
                                torch.as_tensor(np.asarray(samples, dtype=np.float32)[:, :2]))
        return [self.node_list[i] for i in idx.cpu().tolist()]

    def update_maze(self, new_maze):
        self.maze = new_maze
        self.env.maze_map = new_maze
        self._ctx = _ctx_for(np.float32(new_maze), 1.0)

    def check_obstacle_ahead(self, state):
        st = np.asarray(state, dtype=np.float32)[None, :3]
        self._ctx = _ctx_for(self.maze, 1.0)
        return bool(self._ctx.ray_probe(torch.as_tensor(st))[0])

    def extract_path_after_obstacle(self):
        """RRT.py:83-111: the part of the previous main path beyond the first scanned obstacle."""
        path = self.init_main_path[:, :2].copy()
        cur = self.env.state[:2]
        near = int(np.argmin(np.linalg.norm(cur - self.init_main_path[:, :2], axis=1)))
        path = path[near:, :]
        rc = np.array([self.env.cell_xy_to_rowcol(p).astype("int") for p in path])
        hit = -1
        for i, p in enumerate(rc):
            if self.maze[p[0], p[1]] == 1:
                hit = i
                break
        cp = rc[hit]
        while self.maze[cp[0], cp[1]] == 1 and hit < len(rc):
            cp = rc[hit]
            hit += 1
        return path[hit:]

    # ---- planning ---------------------------------------------------------------------------
    def _sample_state(self, remain_init_path):
        if remain_init_path is not None:
            node_idx = np.random.choice(np.arange(len(remain_init_path)))
            return self.random_node_sample() if random.random() < 0.4 else remain_init_path[node_idx][np.newaxis]
        return self.random_node_sample()

    def _pick_goal(self, sample_node):
        if self.run_type == 0:
            return sample_node[0, :2] if random.random() > self.goal_conditioning_bias else self.goal_state[:2]
        return sample_node[0, :2]

    def _sample_batch(self, B):
        """B sampled states and their conditioning goals for one batched expansion round, vectorised: the
        goal-bias coin, the cell draw (run_type >= 2: ONE device inverse-CDF search for all B uniform
        variates) or the uniform position, and the four uniform state components of random_node_sample
        (planners/base_planner.py:162-207); then the goal choice of RRT.py:154-157.  Same distributions as B
        calls of the scalar path, different interleaving of the generator's stream."""
        explore = np.random.random_sample(B) > self.goal_sample_rate
        if self.run_type >= 2:
            pm = np.asarray(self.env.prob_map, dtype=np.float64)
            flat = _ctx_for(self.maze, self.s_global).sample_cells(pm, np.random.random_sample(B)).cpu().numpy()
            rows, cols = np.unravel_index(flat, pm.shape)
            xy = self.env.cell_rowcol_to_xy(np.array([rows, cols])).T
        else:
            xy = np.stack([np.random.uniform(-self.map_width / 2, self.map_width / 2, B),
                           np.random.uniform(-self.map_length / 2, self.map_length / 2, B)], 1)
        rest = np.stack([np.random.uniform(-np.pi, np.pi, B), np.random.uniform(-self.max_v, self.max_v, B),
                         np.random.uniform(-1, 1, B), np.random.uniform(-0.40, 0.40, B)], 1)
        samples = np.where(explore[:, None], np.concatenate([xy, rest], 1), np.asarray(self.goal_state, dtype=np.float64)[None])
        goals = samples[:, :2].copy()
        if self.run_type == 0:
            to_goal = ~(np.random.random_sample(B) > self.goal_conditioning_bias)
            goals[to_goal] = self.goal_state[:2]
        return samples, goals.astype(np.float32)

    def _local_map(self, state):
        n = int(self.local_map_size) if isinstance(self.local_map_size, (int, float)) else int(self.local_map_size[0])
        self._ctx = _ctx_for(self.maze, self.s_global)
        pose = torch.as_tensor(np.asarray(state[:3], dtype=np.float32)[None])
        # the planning loops hand this map to the sampler only, which wants it as 2 m - 1 in bf16 (fm_policy.py:152): the
        # kernel writes that form directly (one launch instead of the map plus three element-wise kernels per iteration)
        # and marks the tensor so that DiffusionSampler.forward takes it as is; common.map_utils.create_local_map is the
        # reference-shaped ({0, 1} float) entry point
        if not getattr(self.sampler, "_accepts_signed_bf16_map", False):
            return self._ctx.local_map(pose, n, self.local_map_scale)   # any other sampler: the reference's {0, 1} map
        lm = self._ctx.local_map(pose, n, self.local_map_scale, bf16_signed=True)
        lm._ditree_signed_bf16 = True
        return lm

    def _plan_device(self):
        """batch_mode = "device": the whole loop on the device (planners/device_planner.py, csrc/planner.cu) for this
        one tree: 256 edge slots, no host work per pass (any run_type, as long as no previous main path is being followed)."""
        from .device_planner import DevicePlanner
        if self.init_main_path is not None:
            raise NotImplementedError("batch_mode='device' does not replan along a previous main path; use 'continuous'")
        start_time = time.time()
        cap = self.iteration_cap if self.iteration_cap is not None else 1 << 20
        cap = max(256, (int(cap) // 256) * 256)
        key = (cap, int(self.action_horizon), tuple(self.prop_duration_schedule), int(self.run_type))
        dp = getattr(self.sampler, "_device_planner", None)
        if dp is None or dp[0] != key or dp[1]._pushed >= dp[1].max_units:
            if dp is not None:
                dp[1].close()
            dp = (key, DevicePlanner(self.sampler, unit_slots=1, iteration_cap=cap, action_horizon=self.action_horizon,
                                     prop_duration=self.prop_duration_schedule, goal_sample_rate=self.goal_sample_rate,
                                     goal_conditioning_bias=self.goal_conditioning_bias,
                                     local_map_scale=self.local_map_scale, max_units=64, run_type=self.run_type))
            self.sampler._device_planner = dp
        unit = dict(start=np.asarray(self.start_node.state, dtype=np.float32), goal=np.asarray(self.env.goal, dtype=np.float32),
                    maze=np.float32(self.maze), maze_name=("maze", self.maze.shape, np.float32(self.maze).tobytes()),
                    seed=int(np.random.randint(0, 2 ** 31 - 1)))
        orig_prob_map = self.env.prob_map.copy() if self.run_type >= 2 else None
        if self.run_type >= 3:
            self.env.update_prob_map_by_loc()              # RRT.py:126-127
        if self.run_type >= 2:
            unit["prob_map"] = np.array(self.env.prob_map, dtype=np.float64)
            unit["prob_key"] = ("pm", unit["prob_map"].tobytes())
            self.env.prob_map = orig_prob_map
        rec = dp[1].run(iter([unit]), time_budget=self.time_budget)[0]
        self.results["iterations"] = rec["results"]["iterations"]
        self.results["number_of_nodes"] = rec["results"]["number_of_nodes"]
        self.results["time"] = time.time() - start_time
        if rec["path"] is None:
            return None, None
        self.results["path"], self.results["actions"] = rec["path"], rec["actions"]
        self.results["path_time"] = len(rec["path"]) * self.env_dt
        return rec["path"], rec["actions"]

    def plan(self):
        if self.batch_size > 1:
            if self.batch_mode == "device":
                return self._plan_device()
            return self._plan_continuous() if self.batch_mode == "continuous" else self._plan_batched()
        start_time = time.time()
        curr_time = time.time()
        total_diffusion_time = 0
        iter_num = 0
        orig_prob_map = self.env.prob_map.copy()
        has_obstacle_ahead = []
        remain_init_path = None
        if self.run_type >= 3:
            self.env.update_prob_map_by_loc()
        if self.run_type > 0 and self.init_main_path is not None:
            remain_init_path = self.extract_path_after_obstacle()
        while (curr_time - start_time) < self.time_budget:
            if self.iteration_cap is not None and iter_num >= self.iteration_cap:
                break
            sample_node = self._sample_state(remain_init_path)
            curr_node = self.nearest_node(sample_node)
            curr_state = curr_node.state
            full_action_seq = None
            full_states_seq = None
            done = False
            prev_actions = curr_node.parent_action_seq
            prev_states = curr_state[None, None, :] if curr_node.parent_states_seq is None else curr_node.parent_states_seq
            edge_length = self.prop_duration_schedule[
                int(np.clip(curr_node.num_visit, 0, len(self.prop_duration_schedule) - 1))]
            curr_node.num_visit += 1
            goal = self._pick_goal(sample_node)
            for _ in range(edge_length // self.action_horizon):
                iter_num += 1
                local_map = self._local_map(curr_state)
                t0 = time.time()
                sampled = self.sampler(prev_states, prev_actions=prev_actions, goal=goal,
                                       local_map=local_map)[0, :self.action_horizon]
                total_diffusion_time += time.time() - t0
                curr_state, done, chunk_actions, chunk_states = self.propagate_action_sequence_env(curr_state, sampled)
                if done is None:  # collision: the whole edge is dropped
                    curr_state = None
                    break
                full_action_seq = chunk_actions if full_action_seq is None else np.concatenate((full_action_seq, chunk_actions))
                prev_actions = chunk_actions
                full_states_seq = chunk_states if full_states_seq is None else \
                    np.concatenate((full_states_seq, chunk_states), axis=1)
                prev_states = chunk_states
                if done:
                    break
            if curr_state is not None and done is not None:
                keep = ~(full_action_seq == 0).all(axis=1)
                full_action_seq = full_action_seq[keep]
                keep = ~(full_states_seq[0] == 0).all(axis=1)
                full_states_seq = full_states_seq[0, keep][np.newaxis]
                new_node = Node(curr_state, full_action_seq, full_states_seq, parent=curr_node)
                self._insert(new_node)
                has_obstacle_ahead.append(False if self.run_type == 0 else self.check_obstacle_ahead(curr_state))
                if done:
                    self.env.prob_map = orig_prob_map
                    return self.handle_goal_reached(new_node, iter_num, start_time)
            curr_time = time.time()
        return self._finish_without_goal(has_obstacle_ahead, iter_num, start_time, orig_prob_map)

    def _finish_without_goal(self, has_obstacle_ahead, iter_num, start_time, orig_prob_map):
        """Budget exhausted (RRT.py:220-257): return the path to the best node."""
        if np.all(has_obstacle_ahead):  # also true for an empty list, like the reference
            self.env.prob_map = orig_prob_map
            self.results["iterations"] = iter_num
            self.results["number_of_nodes"] = len(self.node_list)
            return None, None
        nodes = self.node_list[1:]
        ahead = np.array(has_obstacle_ahead, dtype=bool)
        if self.run_type == 0 or self.init_main_path is None:
            n = len(nodes)
            idx = int(self._ctx.goal_cost_argmin(self._tree.x[1:1 + n], self._tree.y[1:1 + n], self.goal_state[:2],
                                                 torch.as_tensor(ahead))[0])
            best = nodes[idx]
        else:
            along = np.array([int(np.argmin(np.linalg.norm(nd.state[:2] - self.init_main_path[:, :2], axis=1)))
                              if not ahead[i] else -1 for i, nd in enumerate(nodes)])
            best = nodes[int(np.argmax(along))]
        self.env.prob_map = orig_prob_map
        return self.handle_goal_reached(best, iter_num, start_time)

    # ---- batched expansion, continuous refill ---------------------------------------------------
    def _plan_continuous(self):
        """B slots, each running one iteration chain of the reference's loop (sample -> nearest node -> an edge
        of up to edge_length / action_horizon chunks, RRT.py:130-211).  Every device pass advances all B
        slots by one chunk (local map -> conditioning -> sampler -> propagate + collide); a slot whose edge
        ended -- collision: edge dropped (RRT.py:179-184); goal or full length: node inserted (:201-207) --
        is refilled at once, so no slot idles while others finish.  Two slot groups, each with its own CUDA
        stream and its own device context (same packed weights, own activation arena): while one group's pass
        runs the host books the other group's results, draws its replacement samples and finds their nearest
        nodes, and the two passes overlap on the GPU (the persistent GEMMs of one fill the wave tails and
        the small-kernel phases of the other: 1.28x at 256 slots).  One packed host->device and one packed
        device->host copy per pass (pinned buffers)."""
        start_time = time.time()
        B = self.batch_size
        smp = self.sampler
        n_map = int(self.local_map_size) if isinstance(self.local_map_size, (int, float)) else int(self.local_map_size[0])
        ctx = smp._context()
        self._ctx = _ctx_for(self.maze, self.s_global)
        dev = ctx.device
        group_ctx = [ctx, smp._twin_context(B)]
        group_ctx[1].set_map(np.float32(self.maze), self.s_global)   # synchronous; every earlier plan() drained its streams
        h = self.action_horizon
        A = smp.action_dim
        mean_np = smp.metadata["Actions_mean"].astype(np.float32)
        goal_xy = np.asarray(self.env.goal, dtype=np.float64)
        sched = self.prop_duration_schedule
        iter_num = 0
        has_obstacle_ahead = []
        orig_prob_map = self.env.prob_map.copy()
        if self.run_type >= 3:
            self.env.update_prob_map_by_loc()
        W_IN = 6 + A + 2                      # state | previous action | conditioning goal
        W_OUT = h * 6 + h * A + 6 + 2         # trajectory | actions | final state | first_coll, done_step

        class Group:
            pass
        groups = []
        for gi in range(2):
            g = Group()
            g.ctx = group_ctx[gi]
            g.stream = torch.cuda.Stream(device=dev)
            # one pinned block, three contiguous parts (states | previous actions | goals): one H2D copy per
            # pass and the device views need no gather kernels
            g.h_in = torch.empty(B * W_IN, dtype=torch.float32).pin_memory()
            g.hin_s = g.h_in[:B * 6].view(B, 6).numpy()
            g.hin_p = g.h_in[B * 6:B * (6 + A)].view(B, A).numpy()
            g.hin_g = g.h_in[B * (6 + A):].view(B, 2).numpy()
            g.h_out = torch.empty((B, W_OUT), dtype=torch.float32).pin_memory()
            g.parent = [None] * B              # Node the slot's edge grows from
            g.chunk = np.zeros(B, dtype=np.int64)
            g.n_chunks = np.ones(B, dtype=np.int64)
            n_max = max(1, max(sched) // h)
            g.hist_s0 = np.zeros((B, n_max, 6))            # start state of every chunk of the slot's edge
            g.hist_a = np.zeros((B, n_max, h, A))
            g.hist_t = np.zeros((B, n_max, h, 6))
            g.event = torch.cuda.Event()
            g.in_flight = False
            groups.append(g)

        def refill(g, slots):
            k = len(slots)
            if k == 0:
                return
            samples, goals = self._sample_batch(k)
            parents = self.nearest_node_batch(samples)
            hs, hp = g.hin_s, g.hin_p
            for j, b in enumerate(slots):
                p = parents[j]
                g.parent[b] = p
                g.n_chunks[b] = max(1, sched[min(max(p.num_visit, 0), len(sched) - 1)] // h)
                p.num_visit += 1
                g.chunk[b] = 0
                hs[b] = p.state
                hp[b] = mean_np if p.parent_action_seq is None or len(p.parent_action_seq) == 0 \
                    else p.parent_action_seq[-1]
            g.hin_g[slots] = goals

        def launch(g):
            c = g.ctx
            g.stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(g.stream):
                d_in = g.h_in.to(dev, non_blocking=True)
                states, prev = d_in[:B * 6].view(B, 6), d_in[B * 6:B * (6 + A)].view(B, A)
                goals_d = d_in[B * (6 + A):].view(B, 2)
                lm = c.local_map(states, n_map, self.local_map_scale, bf16_signed=True)
                cond = c.build_cond_car(states, prev, goals_d, smp.metadata, float(n_map))
                noise = torch.randn((B, smp.pred_horizon, A), device=dev)
                a = c.fm_sample(noise, cond, lm, smp.num_diffusion_iters, smp.metadata["Actions_mean"],
                                smp.metadata["Actions_std"])
                res = c.propagate_collide(states, a, goal_xy, S=h, want_traj=True)
                packed = torch.cat([res["traj"].reshape(B, h * 6), a[:, :h].reshape(B, h * A), res["final"],
                                    res["first_coll"].float()[:, None], res["done_step"].float()[:, None]], 1)
                g.h_out.copy_(packed, non_blocking=True)
                g.event.record(g.stream)
            g.in_flight = True

        def harvest(g):
            """Book one finished pass of group g; returns the goal node if some slot reached the goal."""
            t_wait = time.perf_counter()
            g.event.synchronize()
            PASS_STATS["gpu_wait_s"] += time.perf_counter() - t_wait
            PASS_STATS["passes"] += 1
            g.in_flight = False
            out = g.h_out.numpy().astype(np.float64)
            traj = out[:, :h * 6].reshape(B, h, 6)
            acts = out[:, h * 6:h * 6 + h * A].reshape(B, h, A)
            fin = out[:, h * 6 + h * A:h * 6 + h * A + 6]
            first = out[:, -2].astype(np.int64)
            done = out[:, -1].astype(np.int64)
            new_nodes, goal_node = [], None
            coll = first >= 0                      # collision: the whole edge is dropped (RRT.py:179-184)
            steps = np.where(done >= 0, done + 1, h)
            ok_slots = np.nonzero(~coll)[0]
            ck = g.chunk[ok_slots]
            g.hist_s0[ok_slots, ck] = g.hin_s[ok_slots]
            g.hist_a[ok_slots, ck] = acts[ok_slots]
            g.hist_t[ok_slots, ck] = traj[ok_slots]
            g.chunk[ok_slots] += 1
            ends = ~coll & ((done >= 0) | (g.chunk >= g.n_chunks))
            goes_on = ~coll & ~ends
            for b in np.nonzero(ends)[0]:
                c, n = int(g.chunk[b]), int(steps[b])   # chunks taken, steps of the last one
                a_seq = g.hist_a[b, :c].reshape(c * h, A)[:(c - 1) * h + n].copy()
                # every chunk starts with its start state, like the reference's states_sequence
                s_all = np.concatenate([g.hist_s0[b, :c, None, :], g.hist_t[b, :c]], axis=1).reshape(c * (h + 1), 6)
                s_seq = s_all[:(c - 1) * (h + 1) + 1 + n].copy()
                node = Node(fin[b].copy(), a_seq, s_seq[None], parent=g.parent[b])
                new_nodes.append(node)
                if done[b] >= 0 and goal_node is None:
                    goal_node = node
            g.hin_s[goes_on] = fin[goes_on]        # the edge goes on from the state it reached
            g.hin_p[goes_on] = acts[goes_on, h - 1]
            free = np.nonzero(coll | ends)[0].tolist()
            self._insert_many(new_nodes)
            if self.run_type == 0:
                has_obstacle_ahead.extend([False] * len(new_nodes))
            else:
                has_obstacle_ahead.extend(self.check_obstacle_ahead(nd.state) for nd in new_nodes)
            return goal_node, free

        try:
            for g in groups:
                refill(g, list(range(B)))
            turn = 0
            while (time.time() - start_time) < self.time_budget:
                if self.iteration_cap is not None and iter_num >= self.iteration_cap:
                    break
                g = groups[turn]
                turn ^= 1
                if g.in_flight:
                    goal_node, free = harvest(g)
                    if goal_node is not None:
                        self.env.prob_map = orig_prob_map
                        return self.handle_goal_reached(goal_node, iter_num, start_time)
                    refill(g, free)
                launch(g)
                iter_num += B
            for g in groups:                          # book what is still in flight
                if g.in_flight:
                    goal_node, _ = harvest(g)
                    if goal_node is not None:
                        self.env.prob_map = orig_prob_map
                        return self.handle_goal_reached(goal_node, iter_num, start_time)
            return self._finish_without_goal(has_obstacle_ahead, iter_num, start_time, orig_prob_map)
        finally:
            # Whatever the exit (goal in either loop, budget, an exception), no pass may still be running on a
            # group's arena, map or pinned buffers when the caller -- or the next plan() -- touches them again.
            for g in groups:
                g.stream.synchronize()

    # ---- batched expansion, one round of edges at a time ------------------------------------------
    def _plan_batched(self):
        from ..expansion import TreeExpander
        start_time = time.time()
        B = self.batch_size
        smp = self.sampler
        n_map = int(self.local_map_size) if isinstance(self.local_map_size, (int, float)) else int(self.local_map_size[0])
        ctx = smp._context()
        self._ctx = _ctx_for(self.maze, self.s_global)
        exp = TreeExpander(ctx, smp.metadata, n_map, self.local_map_scale, num_diffusion_iters=smp.num_diffusion_iters,
                           pred_horizon=smp.pred_horizon, action_horizon=self.action_horizon, action_dim=smp.action_dim)
        mean_np = smp.metadata["Actions_mean"].astype(np.float32)
        iter_num = 0
        has_obstacle_ahead = []
        orig_prob_map = self.env.prob_map.copy()
        if self.run_type >= 3:
            self.env.update_prob_map_by_loc()
        goal_xy = np.asarray(self.env.goal, dtype=np.float64)
        h = self.action_horizon
        while (time.time() - start_time) < self.time_budget:
            if self.iteration_cap is not None and iter_num >= self.iteration_cap:
                break
            samples, goals = self._sample_batch(B)
            parents = self.nearest_node_batch(samples)
            for p in parents:
                p.num_visit += 1
            edge_length = self.prop_duration_schedule[0]
            n_chunks = edge_length // h
            states = torch.as_tensor(np.stack([p.state for p in parents]).astype(np.float32)).to(ctx.device, non_blocking=True)
            # previous action = last action of the parent's edge; roots have none -> the action mean, whose
            # normalised value is 0, which is what the reference feeds for prev_actions=None
            prev_np = np.stack([mean_np if p.parent_action_seq is None or len(p.parent_action_seq) == 0 else
                                np.asarray(p.parent_action_seq[-1], dtype=np.float32) for p in parents])
            prev = torch.as_tensor(prev_np).to(ctx.device, non_blocking=True)
            goals_d = torch.as_tensor(goals).to(ctx.device, non_blocking=True)
            alive = torch.ones(B, dtype=torch.bool, device=ctx.device)
            reached = torch.zeros(B, dtype=torch.bool, device=ctx.device)
            n_iter = torch.zeros((), dtype=torch.int64, device=ctx.device)
            c_s0, c_a, c_t, c_m, c_n = [], [], [], [], []
            # The whole edge is enqueued without a host-device synchronisation: candidates whose edge already
            # ended (collision / goal) ride along masked out, which costs nothing extra on the device (the
            # batch is processed as a whole either way) and lets the host run ahead of the GPU.
            for _ in range(n_chunks):
                n_iter += alive.sum()
                lm = ctx.local_map(states, n_map, self.local_map_scale, bf16_signed=True)
                cond = ctx.build_cond_car(states, prev, goals_d, smp.metadata, float(n_map))
                noise = torch.randn((B, smp.pred_horizon, smp.action_dim), device=ctx.device)
                a = ctx.fm_sample(noise, cond, lm, smp.num_diffusion_iters, smp.metadata["Actions_mean"],
                                  smp.metadata["Actions_std"])
                res = ctx.propagate_collide(states, a, goal_xy, S=h, want_traj=True)
                coll = res["first_coll"] >= 0
                done = res["done_step"] >= 0
                took = alive & ~coll                      # candidates whose chunk is kept
                steps = torch.where(done, res["done_step"] + 1, torch.full_like(res["done_step"], h))
                c_s0.append(states); c_a.append(a[:, :h]); c_t.append(res["traj"]); c_m.append(took); c_n.append(steps)
                reached = reached | (took & done)
                alive = took & ~done                      # a collision drops the whole edge (RRT.py:179-184)
                states = torch.where(took[:, None], res["final"], states)
                prev = torch.where(took[:, None], a[:, h - 1], prev)
            # one round trip per round: five stacked arrays instead of one copy per chunk and tensor
            ok = (alive | reached).cpu().numpy()
            iter_num += int(n_iter.item())
            if ok.any():
                reached_h = reached.cpu().numpy()
                fin = states.cpu().numpy().astype(np.float64)
                S0 = torch.stack(c_s0).cpu().numpy().astype(np.float64)          # (chunks, B, 6)
                A = torch.stack(c_a).cpu().numpy().astype(np.float64)            # (chunks, B, h, 2)
                T = torch.stack(c_t).cpu().numpy().astype(np.float64)            # (chunks, B, h, 6)
                M = torch.stack(c_m).cpu().numpy()
                N = torch.stack(c_n).cpu().numpy()
                new_nodes, goal_node = [], None
                for b in np.nonzero(ok)[0]:
                    a_seq, s_seq = [], []
                    for c in range(n_chunks):
                        if not M[c, b]:
                            break
                        n = int(N[c, b])
                        a_seq.append(A[c, b, :n])
                        s_seq.append(S0[c, b][None])       # every chunk starts with its start state,
                        s_seq.append(T[c, b, :n])          # like the reference's states_sequence
                        if n < h:
                            break
                    node = Node(fin[b], np.concatenate(a_seq), np.concatenate(s_seq)[None], parent=parents[b])
                    new_nodes.append(node)
                    if reached_h[b]:
                        goal_node = node
                        break
                self._insert_many(new_nodes)
                if self.run_type == 0:
                    has_obstacle_ahead.extend([False] * len(new_nodes))
                else:
                    has_obstacle_ahead.extend(self.check_obstacle_ahead(n.state) for n in new_nodes)
                if goal_node is not None:
                    self.env.prob_map = orig_prob_map
                    return self.handle_goal_reached(goal_node, iter_num, start_time)
        return self._finish_without_goal(has_obstacle_ahead, iter_num, start_time, orig_prob_map)
