"""Host side of the device-resident planner (planners/device_planner.py) against a scripted stand-in for libditree:
the feed / poll / fetch logic -- units pulled lazily (so a shared work queue can feed several ranks), at most two
rounds queued ahead, one pass in flight per plan, results returned in pull order -- without a GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from ditreeonlineplanner_b200 import _lib as L
from ditreeonlineplanner_b200.planners import device_planner as dp


class _FakeLib:
    """Each plan: `U` unit slots; a unit finishes after `passes_per_unit` passes; queue semantics of csrc/planner.cu."""

    def __init__(self, passes_per_unit=3):
        self.ppu = passes_per_unit
        self.plans = {}
        self.max_queued_ahead = 0

    def new_plan(self, U):
        h = len(self.plans) + 1
        self.plans[h] = dict(U=U, queue=[], head=0, slots=[None] * U, done=0, passes=0, results={})
        return h

    def _fill(self, p):
        for i in range(p["U"]):
            if p["slots"][i] is None and p["head"] < len(p["queue"]):
                p["slots"][i] = [p["queue"][p["head"]], 0, p["passes"]]
                p["head"] += 1

    def dt_plan_push(self, h, arr, n, stream):
        p = self.plans[h.value]
        for i in range(n):
            p["queue"].append((arr[i].unit_id, arr[i].seed, arr[i].map_slot))
        self.max_queued_ahead = max(self.max_queued_ahead, len(p["queue"]) - p["done"])
        self._fill(p)
        return 0

    def dt_plan_pass(self, h, stream):
        p = self.plans[h.value]
        for i, s in enumerate(p["slots"]):
            if s is None:
                continue
            s[1] += 1
            if s[1] >= self.ppu:
                uid, seed, _ = s[0]
                p["results"][uid] = (seed, s[2], p["passes"])
                p["done"] += 1
                p["slots"][i] = None
        p["passes"] += 1
        self._fill(p)
        p.setdefault("snap", {})[p["passes"] - 1] = (p["head"], len(p["queue"]), p["done"], p["passes"], 0)
        return 0

    def dt_plan_counters(self, h, index, wait, out5):
        snap = self.plans[h.value]["snap"][index]
        for i in range(5):
            out5[i] = snap[i]
        return 0

    def dt_plan_fetch(self, h, uid, hdr, path, acts, cap, stream):
        p = self.plans[h.value]
        r = C.cast(hdr, C.POINTER(L.PlanResult)).contents
        seed, first, last = p["results"][uid]
        r.unit_id, r.goal_reached, r.has_path, r.n_states, r.n_actions = uid, 0, 1, 3, 2
        r.n_nodes, r.iterations, r.first_pass, r.last_pass, r.collisions, r.chunks, r.error = 7, 256 * self.ppu, first, last, 1, 256 * self.ppu, 0
        np.ctypeslib.as_array(C.cast(path, C.POINTER(C.c_float)), (3, 6))[:] = seed % 1000
        np.ctypeslib.as_array(C.cast(acts, C.POINTER(C.c_float)), (2, 2))[:] = uid
        return 0


class _FakeCtx:
    device = torch.device("cpu")

    def __init__(self):
        self.slots = {}

    def _check(self, rc):
        assert rc == 0

    def _stream(self):
        return C.c_void_p(0)

    def set_map_slot(self, slot, grid, s):
        self.slots[slot] = grid.shape

    def sync_status(self):
        pass


def _planner(lib, U, streams):
    pl = dp.DevicePlanner.__new__(dp.DevicePlanner)
    pl.lib, pl.U, pl.max_units, pl.max_path, pl.stats, pl._pushed, pl.run_type = lib, U, 1000, 16, {}, 0, 0
    pl.plans = []
    for i in range(streams):
        p = dp._Plan()
        p.U = U // streams + (1 if i < U % streams else 0)
        p.ctx, p.stream = _FakeCtx(), None
        p.h = C.c_void_p(lib.new_plan(p.U))
        p.pushed = p.done = p.passes = 0
        p.maps, p.cdfs, p.order = {}, {}, []
        pl.plans.append(p)
    return pl


@pytest.mark.parametrize("U,streams,n_units", [(4, 1, 19), (8, 2, 37), (6, 2, 5), (3, 1, 0)])
def test_feed_poll_fetch(U, streams, n_units):
    lib = _FakeLib(passes_per_unit=3)
    pl = _planner(lib, U, streams)
    pulled = []

    def source():
        for i in range(n_units):
            pulled.append(i)
            yield dict(start=np.zeros(6, np.float32), goal=np.ones(2, np.float32), maze=np.zeros((5 + i % 2, 7)),
                       maze_name=f"m{i % 2}", seed=1000 + i)

    recs = pl.run(source())
    assert len(recs) == n_units and pulled == list(range(n_units))
    for i, r in enumerate(recs):               # pull order, whatever plan a unit went to
        assert r["finished"] and r["path"].shape == (3, 6) and r["actions"].shape == (2, 2)
        assert r["path"][0, 0] == (1000 + i) % 1000
        assert r["results"] == {"iterations": 768, "number_of_nodes": 7, "path_time": 3 * 0.02}
        assert r["runtime"] >= 0.0
    # never more than two rounds queued ahead of what has finished (so a shared queue stays shared)
    assert lib.max_queued_ahead <= 2 * max(p.U for p in pl.plans)
    # both mazes staged once per plan that saw them
    for p in pl.plans:
        assert set(p.ctx.slots.values()) <= {(5, 7), (6, 7)}
    if n_units:
        assert pl.stats["units"] == n_units and pl.stats["passes"] >= 3


def test_lazy_pull_leaves_units_for_other_ranks():
    """A second consumer of the same iterator (another rank's planner on the shared counter) still finds units."""
    lib = _FakeLib(passes_per_unit=2)
    pl = _planner(lib, 2, 1)
    it = iter([dict(start=np.zeros(6, np.float32), goal=np.ones(2, np.float32), maze=np.zeros((5, 5)), maze_name="m",
                    seed=i) for i in range(40)])
    taken = []

    def source():
        for u in it:
            taken.append(u["seed"])
            yield u
            if len(taken) == 6:
                # "another rank" drains the rest of the shared queue meanwhile
                list(it)
    recs = pl.run(source())
    assert len(recs) == 6 == len(taken)
