"""Device-resident, multi-scenario RRT: the host side of ``dt_plan_*`` (csrc/planner.cu).

The reference plans one scenario at a time, one sampler call per loop iteration (planners/RRT.py:113-257 inside
run_scenarios.py:202-395).  Here ``unit_slots`` (scenario, run) units grow their trees concurrently in one device
pass -- 256 edges each, so a pass is ``unit_slots * 256`` candidates through local map -> sampler -> propagation ->
collision -- and sampling, nearest-node search, node insertion, goal / budget tests, the final node selection and the
path back-trace all stay on the device.  The host enqueues passes, feeds unit descriptors to the device queue and
reads five counters with a lag of one pass; results are fetched once, at the end.

Semantics are those of ``RRT_Planner(batch_size=256, batch_mode="continuous")`` (each pass advances every edge by one
chunk and refills finished slots at once; the samples of a pass see the tree as of the start of the pass) with
``run_type = 0``; random numbers come from per-unit Philox streams on the device, so a unit's result depends on its
seed only -- not on the GPU count, the rank that ran it or the units that shared its passes.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from .. import _lib as L

EDGE_SLOTS = 256


class _Plan:
    """One dt_plan on one device context and stream."""
    first_done = 0
    first_done_total = 0


class DevicePlanner:
    def __init__(self, sampler, unit_slots=8, iteration_cap=4096, action_horizon=8, prop_duration=(64,),
                 goal_sample_rate=0.15, goal_conditioning_bias=0.85, local_map_scale=0.2, max_units=256, max_path=4096,
                 streams=1, run_type=0):
        """`sampler`: the DiffusionSampler mirror that owns the packed denoiser (its max_batch should be >=
        unit_slots * 256 / streams); the remaining arguments are RRT_Planner's (planners/RRT.py:19-31).
        streams = 2 splits the unit slots over two plans on two CUDA streams and two device contexts (same packed
        weights, own activation arenas): the short kernels at the start and end of one plan's pass (sampling, nearest
        node, local maps, the first encoder layers, insertion) overlap the other plan's tensor-core phase."""
        self.sampler = sampler
        self.ctx = sampler._context()
        self.lib = self.ctx.lib
        self.U = int(unit_slots)
        self.iteration_cap = int(iteration_cap)
        self.max_units = int(max_units)
        self.max_path = int(max_path)
        streams = max(1, min(int(streams), self.U))
        self.run_type = int(run_type)
        cfg = L.PlanCfg()
        cfg.run_type = self.run_type
        cfg.edge_slots = EDGE_SLOTS
        cfg.node_cap = self.iteration_cap + 1           # every node costs at least one chunk expansion
        cfg.action_horizon = int(action_horizon)
        sched = [max(1, int(d) // int(action_horizon)) for d in prop_duration][:8]
        cfg.n_sched = len(sched)
        for i, v in enumerate(sched):
            cfg.sched_chunks[i] = v
        cfg.iteration_cap = self.iteration_cap
        cfg.ode_steps = int(sampler.num_diffusion_iters)
        cfg.max_units, cfg.max_path = self.max_units, self.max_path
        cfg.goal_sample_rate, cfg.goal_conditioning_bias = float(goal_sample_rate), float(goal_conditioning_bias)
        cfg.local_map_scale = float(local_map_scale)
        for i, v in enumerate(self.ctx._norm_vec(sampler.metadata)):
            cfg.norm[i] = float(v)
        self.plans = []
        for i in range(streams):
            pl = _Plan()
            pl.U = self.U // streams + (1 if i < self.U % streams else 0)
            pl.ctx = self.ctx if i == 0 else sampler._twin_context(pl.U * EDGE_SLOTS)
            pl.stream = None if streams == 1 else torch.cuda.Stream(device=self.ctx.device)
            cfg.unit_slots = pl.U
            h = C.c_void_p()
            pl.ctx._check(self.lib.dt_plan_create(pl.ctx.h, C.byref(cfg), C.byref(h)))
            pl.h = h
            pl.pushed = pl.done = pl.passes = 0
            pl.maps = {}
            pl.cdfs = {}
            pl.order = []
            self.plans.append(pl)
        self.h = self.plans[0].h
        self._pushed = 0
        self.stats = {}

    def close(self):
        for pl in getattr(self, "plans", []):
            if pl.h:
                self.lib.dt_plan_destroy(pl.h)
                pl.h = None
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- maps ----------------------------------------------------------------------------------------------
    def _map_slot(self, pl, name, grid):
        """Slot of maze `name` in plan pl's context, staging it on first use (dt_set_map_slot)."""
        hit = pl.maps.get(name)
        if hit is None:
            slot = len(pl.maps)
            g = np.asarray(grid, dtype=np.float32)
            pl.ctx.set_map_slot(slot, g, 1.0)
            hit = pl.maps[name] = (slot, g.shape[0], g.shape[1])
        return hit

    def _cdf_slot(self, pl, key, prob_map):
        """Slot of a unit's probability map (CarEnv.prob_map) in plan pl, staged on first use (dt_plan_set_cdf)."""
        hit = pl.cdfs.get(key)
        if hit is None:
            if len(pl.cdfs) >= 64:
                raise RuntimeError("DevicePlanner: more than 64 distinct probability maps")
            hit = pl.cdfs[key] = len(pl.cdfs)
            pm = np.ascontiguousarray(np.asarray(prob_map, dtype=np.float64).ravel())
            with self._stream_of(pl):
                pl.ctx._check(self.lib.dt_plan_set_cdf(pl.h, hit, pm.ctypes.data_as(C.c_void_p), pm.size, pl.ctx._stream()))
        return hit

    def _stream_of(self, pl):
        return torch.cuda.stream(pl.stream) if pl.stream is not None else _NullCtx()

    # ---- the loop ------------------------------------------------------------------------------------------
    def _push(self, units, plan=None):
        pl = self.plans[0] if plan is None else plan
        arr = (L.PlanUnit * len(units))()
        for i, u in enumerate(units):
            slot, rows, cols = self._map_slot(pl, u["maze_name"], u["maze"])
            for k in range(6):
                arr[i].start[k] = float(u["start"][k])
            arr[i].goal[0], arr[i].goal[1] = float(u["goal"][0]), float(u["goal"][1])
            arr[i].half_w, arr[i].half_h = cols / 2.0, rows / 2.0   # map_width = len(maze[0]), map_length = len(maze)
            arr[i].map_slot = slot
            arr[i].seed = int(u["seed"]) & 0xFFFFFFFF
            arr[i].unit_id = pl.pushed + i
            arr[i].cdf_slot = -1
            if self.run_type >= 2:
                if u.get("prob_map") is None:
                    raise ValueError("run_type >= 2 needs the unit's prob_map (CarEnv(run_type).prob_map)")
                arr[i].cdf_slot = self._cdf_slot(pl, u.get("prob_key", (u["maze_name"],)), u["prob_map"])
        with self._stream_of(pl):
            pl.ctx._check(self.lib.dt_plan_push(pl.h, arr, len(units), pl.ctx._stream()))
        pl.pushed += len(units)
        self._pushed += len(units)

    def run(self, unit_source, time_budget=None):
        """Plan every unit `unit_source` yields (dicts with start (6,), goal (2,), maze (R,C), maze_name, seed; pulled
        lazily, so the source may be a work queue shared between ranks).  -> list of result dicts in pull order:
        path (n,6) | None, actions (m,2) | None, results {iterations, number_of_nodes, path_time}, runtime, goal_reached,
        collisions, chunks."""
        it = iter(unit_source)
        exhausted = False
        t_start = time.perf_counter()
        wait_s = 0.0
        counters = (C.c_int32 * 5)()
        pull_order = []            # (plan index, unit id inside that plan) in pull order
        for pl in self.plans:
            pl.first = pl.pushed
            pl.base_pass = pl.passes
            pl.n_pass = 0
            pl.done = 0
        pass_times = [dict() for _ in self.plans]   # per plan: pass index -> host time at which it was seen complete

        def feed():
            nonlocal exhausted
            for pi, pl in enumerate(self.plans):
                batch = []
                while not exhausted and (pl.pushed + len(batch)) - (pl.first + pl.done) < 2 * pl.U:
                    try:
                        unit = next(it)
                    except StopIteration:
                        exhausted = True
                        break
                    if pl.pushed + len(batch) >= self.max_units:
                        raise RuntimeError(f"DevicePlanner: more than max_units = {self.max_units} units per plan")
                    batch.append(unit)
                    pull_order.append((pi, pl.pushed + len(batch) - 1))
                if batch:
                    self._push(batch, pl)

        def poll(pi, pl, index):
            nonlocal wait_s
            t0 = time.perf_counter()
            pl.ctx._check(self.lib.dt_plan_counters(pl.h, index, 1, counters))
            now = time.perf_counter()
            wait_s += now - t0
            pass_times[pi][index] = now
            pl.done = int(counters[2]) - pl.first_done
            if counters[4] & 1:
                raise RuntimeError("DevicePlanner: node capacity exhausted")

        for pl in self.plans:
            pl.first_done = getattr(pl, "first_done_total", 0)
        feed()
        while True:
            pending = [pl for pl in self.plans if not (exhausted and pl.done >= pl.pushed - pl.first)]
            if not pending:
                break
            if time_budget is not None and time.perf_counter() - t_start > time_budget:
                break
            for pi, pl in enumerate(self.plans):
                if exhausted and pl.done >= pl.pushed - pl.first:
                    continue
                with self._stream_of(pl):
                    pl.ctx._check(self.lib.dt_plan_pass(pl.h, pl.ctx._stream()))
                pl.n_pass += 1
            for pi, pl in enumerate(self.plans):
                if pl.n_pass >= 2 and (pl.base_pass + pl.n_pass - 2) not in pass_times[pi]:
                    poll(pi, pl, pl.base_pass + pl.n_pass - 2)   # one pass stays in flight per plan
            feed()
        for pi, pl in enumerate(self.plans):
            if pl.n_pass:
                poll(pi, pl, pl.base_pass + pl.n_pass - 1)
            pl.passes = pl.base_pass + pl.n_pass
            pl.first_done_total = getattr(pl, "first_done_total", 0) + pl.done
            with self._stream_of(pl):
                pl.ctx.sync_status()
        wall = time.perf_counter() - t_start
        # results, fetched once
        out = []
        hdr = L.PlanResult()
        path = np.empty((self.max_path, 6), np.float32)
        acts = np.empty((self.max_path, 2), np.float32)
        for pi, uid in pull_order:
            pl = self.plans[pi]
            with self._stream_of(pl):
                pl.ctx._check(self.lib.dt_plan_fetch(pl.h, uid, C.byref(hdr), path.ctypes.data_as(C.c_void_p),
                                                     acts.ctypes.data_as(C.c_void_p), self.max_path, pl.ctx._stream()))
            finished = hdr.unit_id == uid
            t1 = pass_times[pi].get(hdr.last_pass, t_start + wall)
            t0 = pass_times[pi].get(hdr.first_pass - 1, t_start)
            res = {"iterations": int(hdr.iterations) if finished else 0, "number_of_nodes": int(hdr.n_nodes) if finished else 0}
            rec = dict(path=None, actions=None, results=res, runtime=max(t1 - t0, 0.0), finished=bool(finished),
                       goal_reached=bool(hdr.goal_reached) if finished else False,
                       collisions=int(hdr.collisions) if finished else 0, chunks=int(hdr.chunks) if finished else 0,
                       error=int(hdr.error) if finished else 0)
            if finished and hdr.has_path:
                rec["path"] = path[:hdr.n_states].copy()
                rec["actions"] = acts[:hdr.n_actions].copy()
                res["path_time"] = hdr.n_states * 0.02      # len(path) * env.dt (base_planner.py:237)
            out.append(rec)
        self.stats = dict(passes=sum(pl.n_pass for pl in self.plans), wall_s=wall, device_wait_s=wait_s, units=len(out),
                          candidates_per_pass=self.U * EDGE_SLOTS, streams=len(self.plans))
        return out


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
